set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_final_2gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_final_2gpu.log
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench n2 rc=$? elapsed ${SECONDS}s"
tail -c 1500 gpurun_out/r02_bench_n2_final.json
