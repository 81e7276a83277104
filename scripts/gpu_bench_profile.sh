# one GPU: bench (both arms), then the ncu launch list and one full capture of the megakernel
set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
cat gpurun_out/bench_ref.json gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out
