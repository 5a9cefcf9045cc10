# round-end check on one B200: GPU tests, smoke, the default bench line, one A/B, ncu launch list + full capture of one step
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_final.log
/usr/bin/time -v python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; grep "Elapsed" gpurun_out/r02_bench_final.err
NGCF_B200_WGRAD_OVERLAP=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-epoch --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wgrad overlap=1: step', d['ms_per_step'], 'warm', d['warm_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch --no-extra --eager > /dev/null 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spmm_stream|compact_kernel|dense_fwd_tma|dense_bwd_tc|wgrad_tc|hub_finish" -s 69 -c 23 -f -o gpurun_out/r02_final_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch --no-extra --eager > /dev/null 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02_final_step.ncu-rep --page raw --csv > gpurun_out/r02_final_step_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
