# compute-sanitizer passes over every kernel family (run through gpurun on one B200); logs under gpurun_out/sanitize_*.log
# NOTE: this pod's gpurun refuses compute-sanitizer ("closed on this pool"), so the passes could not be run here; the script
# is kept for boxes that allow it.  scripts/sanitize_target.py alone is a plain every-kernel-family smoke run.
# usage: bash scripts/sanitize.sh [tools...]   (default: memcheck racecheck synccheck initcheck)
tools=${@:-memcheck racecheck synccheck initcheck}
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for t in $tools; do
  timeout 900 compute-sanitizer --tool $t --error-exitcode 9 --print-limit 30 python scripts/sanitize_target.py > gpurun_out/sanitize_$t.log 2>&1
  echo "$t target rc=$? $(grep -c 'Invalid\|Race\|hazard\|Uninitialized\|Barrier error' gpurun_out/sanitize_$t.log) report lines; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitize_$t.log | tail -1)"
done
# the time-chunked path (chunk_run, p2a / deemph / renorm chunk calls) through its own test, memcheck only
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_griffinlim.py -x -q -m gpu -k "emulated and 449 or chunk_step or windowed_carry" > gpurun_out/sanitize_memcheck_chunk.log 2>&1
echo "memcheck chunk tests rc=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_chunk.log | tail -1) $(tail -1 gpurun_out/sanitize_memcheck_chunk.log)"
