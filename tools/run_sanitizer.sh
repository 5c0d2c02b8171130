# one compute-sanitizer tool per gpurun call:  bash tools/run_sanitizer.sh memcheck|racecheck|synccheck|initcheck [cases...]
TOOL=${1:-memcheck}; shift
timeout 300 python tools/sanitize_cases.py "$@" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_cases.py "$@" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "sanitizer exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|items" gpurun_out/sanitize_$TOOL.log | tail -20
