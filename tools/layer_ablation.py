"""Time single conv layers (gd_debug_tapgemm) at benchmark scale under the GDECONV_ABL ablation switches."""
import ctypes as C, os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'galaxy-deconv_b200'), ROOT]
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    from gdeconv._lib import lib, check
    dev = torch.device('cuda:0')
    out = {}
    for name, H, ntaps, Kt, N, batch in [('L0', 48, 9, 32, 32, 4096), ('L1', 24, 9, 64, 64, 4096), ('L2', 12, 9, 128, 128, 4096), ('L3', 6, 9, 256, 256, 4096),
                                          ('down0', 24, 1, 128, 64, 4096), ('up1', 12, 1, 128, 256, 4096)]:
        g = (C.c_int * 7)(); check(lib.gd_debug_geom(H, batch, C.byref(g))); Ptot = g[5]
        act = torch.zeros(Kt // 8, Ptot, 8, dtype=torch.float16, device=dev).normal_()
        w = (torch.randn(ntaps, Kt // 8, N, 8, device=dev) * 0.05).half()
        o = torch.empty(N // 4, Ptot, 4, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        f = lambda: check(lib.gd_debug_tapgemm(1, H, batch, ntaps, Kt, N, 0, C.c_void_p(act.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(o.data_ptr()), st))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        fl = 2.0 * batch * H * H * N * Kt * ntaps
        out[name] = (round(us, 1), round(fl / us / 1e6, 1))
    print(json.dumps(out))
else:
    for abl in (0, 1, 2, 3, 4, 7):
        env = dict(os.environ, GDECONV_ABL=str(abl))
        r = subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True)
        print('ABL', abl, r.stdout.strip() or r.stderr[-500:])
