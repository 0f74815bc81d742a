for v in "" "GDECONV_CHUNK=10240" "GDECONV_CHUNK=10240 GDECONV_STREAMS=1" "GDECONV_CHUNK=3584" "GDECONV_CHUNK=2560" ""; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', round(d['value']), round(d['e2e']['value']), d['config']['chunk'], d['config']['streams'])"
done
