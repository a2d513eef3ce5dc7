#!/bin/bash
# Regenerates every measured artefact under profiles/ on a B200 box (one GPU):
#   gpurun --timeout 3000 -- 'bash scripts/gpu_evidence.sh r2'
# Bench lines of BASELINE.json configs 1-5, the ncu launch list, the --set full captures (details + DRAM traffic), the
# per-page-size DRAM counters, the sanitizer logs. Everything lands in gpurun_out/ev_<tag>_*; scripts/collect_profiles.py
# turns the captures into the committed summaries. Numbers printed by runs under ncu / compute-sanitizer are never bench values.
TAG=${1:-r2}
O=gpurun_out/ev_${TAG}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > ${O}_pytest.log
# configs[3] (default: 16 GiB mixed, the headline), configs[1] (random), configs[2] (textures), the reference arm
python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.err
python bench.py --impl reference > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err
python bench.py --workload random --no-secondary > ${O}_bench_random.json 2> ${O}_bench_random.err
python bench.py --workload texture --no-secondary > ${O}_bench_texture.json 2> ${O}_bench_texture.err
# configs[0]: one 64 KiB page / 1 MiB of the low-entropy source, CPU DecodeCPU vs the kernel, bit-exact
timeout 300 python scripts/config0_check.py > ${O}_config0.log 2>&1
# configs[4]: page-size sweep, 8 GiB per size, then one launch per size under ncu for the DRAM counters
timeout 1500 python scripts/page_size_sweep.py 8192 > ${O}_sweep.log 2>&1; cp gpurun_out/page_size_sweep.json ${O}_page_size_sweep.json
SWEEP_SINGLE=1 timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:bgx_decode_pages --csv --log-file ${O}_sweep_dram.csv python scripts/page_size_sweep.py 2048 > ${O}_sweep_ncu.log 2>&1
timeout 300 python scripts/gpu_bench_kinds.py 16 32 text,binary,mixed,lowent,texture > ${O}_kinds.log 2>&1
# launch list of the bench command (kernel shares of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --size-gib 2 --no-cpu --no-e2e --no-secondary > ${O}_launches.log 2>&1
# --set full captures of the page kernel: mixed (the headline payload), raw pages, textures
ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -s 1 -c 1 -f -o ${O}_mixed python scripts/gpu_prof_one.py mixed 64 16 2 > ${O}_ncu_mixed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -s 1 -c 1 -f -o ${O}_raw python scripts/gpu_prof_one.py random 64 16 2 > ${O}_ncu_raw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgx_de -s 2 -c 2 -f -o ${O}_texture python scripts/gpu_prof_one.py texture 16 64 2 > ${O}_ncu_texture.log 2>&1
# sanitizers on a small decode (raw, compressed, long-run and partial pages)
compute-sanitizer --tool memcheck python scripts/gpu_sanity_small.py 2>&1 | tail -4 > ${O}_sanitizer_memcheck.log
compute-sanitizer --tool synccheck python scripts/gpu_sanity_small.py 2>&1 | tail -4 > ${O}_sanitizer_synccheck.log
cat ${O}_pytest.log; tail -c 400 ${O}_bench_n1.json; tail -3 ${O}_sweep.log
# lock-step verification build under racecheck (every hand-over a CTA barrier, piece loads byte-exact): the hazards of the
# production build that are ordered by mbarriers or are discarded over-reads must all disappear
if [ -f build/variants/libbgx_sync.so ]; then
  BGX_CUDA_LIB=$PWD/build/variants/libbgx_sync.so timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/gpu_sanity_small.py 2>&1 | tail -6 > ${O}_sanitizer_racecheck_lockstep.log
fi
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/gpu_sanity_small.py 2>&1 | grep -E "RACECHECK SUMMARY|SANITY" > ${O}_sanitizer_racecheck_production.log
# where a 4 KiB page spends its time: one --set full capture of the 4 KiB single-page-stream launch
SWEEP_SINGLE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -c 1 -f -o ${O}_page4k python scripts/page_size_sweep.py 1024 mixed 4096 > ${O}_ncu_page4k.log 2>&1
