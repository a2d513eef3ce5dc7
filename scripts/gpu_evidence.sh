set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/ev_pytest.log
python bench.py > gpurun_out/ev_bench_n1.json 2> gpurun_out/ev_bench_n1.err
python bench.py --workload texture --size-gib 2 --no-secondary > gpurun_out/ev_bench_texture.json 2> gpurun_out/ev_bench_texture.err
python bench.py --workload mixed --size-gib 4 --no-secondary --no-cpu > gpurun_out/ev_bench_mixed4g.json 2> gpurun_out/ev_bench_mixed4g.err
timeout 300 python scripts/page_size_sweep.py 2048 > gpurun_out/ev_sweep.log 2>&1
timeout 200 python scripts/gpu_bench_kinds.py > gpurun_out/ev_kinds.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 2 --warmup 3 --size-gib 1 --no-cpu --no-e2e > gpurun_out/ev_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -s 1 -c 1 -f -o gpurun_out/ev_raw python scripts/gpu_prof_one.py random 64 16 2 > gpurun_out/ev_ncu_raw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -s 1 -c 1 -f -o gpurun_out/ev_mixed python scripts/gpu_prof_one.py mixed 64 16 2 > gpurun_out/ev_ncu_mixed.log 2>&1
cat gpurun_out/ev_pytest.log; tail -c 600 gpurun_out/ev_bench_texture.json; tail -3 gpurun_out/ev_sweep.log
