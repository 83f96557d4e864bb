#!/bin/bash
# Round-2 evidence, one GPU: bench lines of every workload (with clocks, roofline, cpu_baseline), the reference arm,
# ncu launch lists and --set full captures of the cutoff-path kernels and of the FP64 peak probe.  Run under gpurun
# from the repository root; everything lands in gpurun_out/, profiles/summarize.py condenses it.
set -u
out=gpurun_out
python bench.py --steps 20 --warmup 5 > $out/bench_r02_ions1m.json 2> $out/bench_r02_ions1m.err
python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_r02_ions1m_reference.json 2> $out/bench_r02_ions1m_reference.err
for w in water3k solvated25k opt5k; do
  python bench.py --workload $w --steps 50 --warmup 5 > $out/bench_r02_$w.json 2> $out/bench_r02_$w.err
done
# launch lists (each after the same command exited 0 without ncu)
for w in ions1m solvated25k water3k; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $out/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 90 --csv --log-file $out/launches_r02_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$w.log 2>&1
done
# --set full: the three kernels of the cluster scheme (one launch each) and the DFMA peak probe
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cl_pair|cl_list|cl_reduce' -s 9 -c 3 -o $out/prof_r02_cluster \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:dfma_peak -c 1 -o $out/prof_r02_dfma_peak \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_peak.log 2>&1
tail -c 300 $out/bench_r02_ions1m.json
