""""reference eager PyTorch on the same B200" as the GPU-LIBRARY baseline (cuFFT + cuDNN, the thing a
user of the reference gets by moving its nn.Module to cuda:0).  The oracle port of UnrolledADMMGaussian(8) (plain torch,
bit-exact vs the reference on CPU) is moved to the GPU and timed with CUDA events on the same synthetic stamps, batch 1024,
in fp32 with TF32 off and on (torch's cuDNN default).  Measurement only: nothing of the product path runs here.
    python tools/eager_gpu_baseline.py > profiles/eager_gpu_r01.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]
import torch
import oracle.ref_models as O
from gdeconv.synth import make_batch

dev = torch.device('cuda:0')
N, B = 4096, 1024
data = make_batch(0, N, 100.0, device=dev)
m = O.UnrolledADMMGaussian(8).eval()
m.load_state_dict(O.seeded_state_dict(lambda: O.UnrolledADMMGaussian(8), 12))
m = m.to(dev)
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        def run():
            for i in range(0, N, B):
                m(data['obs'][i:i + B], data['psf'][i:i + B], data['alpha'][i:i + B])
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps(dict(model='UnrolledADMMGaussian(8) oracle port, eager torch %s on %s' % (torch.__version__, torch.cuda.get_device_name(0)),
                          stamps=N, batch=B, tf32=tf32, cudnn_benchmark=True, ms_per_pass=ms, galaxies_per_s=N / (ms * 1e-3))), flush=True)
