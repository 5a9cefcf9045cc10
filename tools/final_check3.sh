set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_final.log
SECONDS=0
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? elapsed ${SECONDS}s"
