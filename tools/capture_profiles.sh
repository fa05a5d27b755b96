#!/bin/bash
# Round-2 profile captures (run under gpurun on one B200; every ncu run follows a plain run of the same command).
#   bash tools/capture_profiles.sh        -> gpurun_out/r02_*.ncu-rep, gpurun_out/r02_bench_launches.csv
# Digest here afterwards with tools/ncu_summary.py / tools/ncu_exec.py / tools/summarize_launches.py (see profiles/README.md).
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
python tools/t_one.py feat        > $O/p1.log 2>&1 && $NCU -k regex:fused_mg         -o $O/r02_mg_feat      python tools/t_one.py feat        > $O/n1.log 2>&1
python tools/t_one.py rays        > $O/p2.log 2>&1 && $NCU -k regex:fused_mg         -o $O/r02_mg_rays      python tools/t_one.py rays        > $O/n2.log 2>&1
python tools/t_one.py feat 32768  > $O/p3.log 2>&1 && $NCU -k regex:fused_mg         -o $O/r02_mg_feat_big  python tools/t_one.py feat 32768  > $O/n3.log 2>&1
python tools/t_one_f32.py         > $O/p4.log 2>&1 && $NCU -k regex:fused_f32_kernel -o $O/r02_f32          python tools/t_one_f32.py         > $O/n4.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render --no-extra --eager > $O/p5.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_bench_launches.csv \
      python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render --no-extra --eager > $O/n5.log 2>&1
for f in p1 p2 p3 p4; do tail -n 1 $O/$f.log; done
python tools/t_render.py 8        > $O/p6.log 2>&1 && $NCU -k regex:fused_v1         -o $O/r02_render_fwd   python tools/t_render.py 2        > $O/n6.log 2>&1
tail -n 1 $O/p6.log
