# usage: bash tools/run_variants.sh <variant names...>  (libraries under variants/lib_<name>.so)
python -m pytest tests -m gpu -x -q -k "cell or cutoff or config3 or config5 or pair" 2>&1 | tail -3
for v in "$@"; do
  for wl in ions1m solvated25k; do
    JME_B200_LIB=$PWD/variants/lib_$v.so python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | V=$v WL=$wl python -c "
import json,sys,os
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('%s %s kernel_ms %.3f build_ms %.3f ms_step %.3f e2e_ms %.3f value %.3e'%(os.environ['V'],os.environ['WL'],r['kernel_ms'],r['build_ms'],d['ms_per_step'],d['e2e']['ms_per_step'],d['value']))"
  done
done
