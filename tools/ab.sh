#!/bin/bash
# development helper (GPU box): bench.py once per library variant under variants/ (and the in-tree library as "main")
for v in main "$@"; do
  if [ "$v" = main ]; then unset JME_B200_LIB; else export JME_B200_LIB=$PWD/variants/lib_$v.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['roofline']['build_ms'],4))"
done
