set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3d_bench_ref.json 2>> gpurun_out/r3d_bench.err
python tools/microbench.py > gpurun_out/r3d_microbench.log 2>&1
for sm in 0 2097152 4194304 8388608; do echo "stream_min=$sm"; python tools/sweep.py --host --sizes 20,21,22,23,24 --dists uniform --steps 3 --stream-min $sm 2>&1 | grep -v "^\["; done > gpurun_out/r3d_e2e_sweep.log 2>&1
python tools/sweep.py --sizes 16,18,20,22,24,26 --dists uniform,const,wminus --steps 3 2>&1 | grep -v "^\[" > gpurun_out/r3d_sweep.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 10 -c 4 -o gpurun_out/r3d_prof_accumulate -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3d_ncu2.log 2>&1
ls -la gpurun_out/ | tail -15
