set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_h.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_h.log
for pf in 0 2 4; do
  echo "== PREFETCH=$pf"
  NGCF_B200_PREFETCH=$pf python tools/kernel_times.py 2>/dev/null | grep -v "FFMA\|Philox\|bits" 
  NGCF_B200_PREFETCH=$pf python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-epoch --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'warm', d['warm_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spmm0', d['roofline']['kernel_ms'])"
done
NGCF_B200_PREFETCH=0 python tools/bwd_timeline.py 2>/dev/null | tail -22
NGCF_B200_PREFETCH=2 python tools/bwd_timeline.py 2>/dev/null | tail -22
