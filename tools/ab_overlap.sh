set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_tests_final.log
run() { NGCF_B200_COMPACT_OVERLAP=$1 NGCF_B200_FEATMIX_OVERLAP=$2 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-epoch --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('compact=$1 featmix=$2: step', d['ms_per_step'], 'warm', d['warm_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches_per_step'], 'loss', d['e2e']['last_loss'])"; }
run 0 0; run late 1; run 0 0; run late 1; run late 0; run 0 1
