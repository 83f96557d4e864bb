for v in exp_NOMATH exp_NOGATHER exp_BOTH; do
    JME_B200_LIB=$PWD/variants/lib_$v.so python bench.py --workload ions1m --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | V=$v python -c "
import json,sys,os
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('%s kernel_ms %.3f build_ms %.3f ms_step %.3f'%(os.environ['V'],r['kernel_ms'],r['build_ms'],d['ms_per_step']))"
done
