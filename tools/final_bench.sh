set -x
SECONDS=0
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? elapsed ${SECONDS}s"
for ov in 0 1 0 1; do
NGCF_B200_WGRAD_OVERLAP=$ov python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-epoch --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wgrad overlap=$ov: step', d['ms_per_step'], 'warm', d['warm_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
