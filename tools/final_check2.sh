set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_tests_final.log
SECONDS=0
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? elapsed ${SECONDS}s"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm_|compact_kernel|dense_|wgrad_|hub_finish|featmix|gather_concat|bpr_kernel|rowgrad|pack_weights" -c 140 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch --no-extra --eager > /dev/null 2>&1; echo "ncu list rc=$?"
