#!/bin/sh
# compute-sanitizer passes over the hot path (the reference has no race detection; the new build
# checks its kernels).  Runs __graft_entry__.smoke() -- one small mean-teacher iteration through every kernel of the
# default path, eager launches -- under memcheck, racecheck (shared-memory hazards of the TMA / mbarrier pipelines) and
# synccheck.  Needs a B200; each pass takes minutes (the tools serialise and instrument every launch), so give gpurun a
# generous --timeout.  Reports land in gpurun_out/sanitize_<tool>.log; the exit code is the number of failing tools.
#   sh tools/sanitize.sh [memcheck racecheck synccheck]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${*:-"memcheck racecheck synccheck"}
fail=0
for tool in $TOOLS; do
    log=gpurun_out/sanitize_$tool.log
    DCASE_NO_GRAPH=1 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -c "import __graft_entry__ as g; g.smoke()" > "$log" 2>&1
    rc=$?
    tail -3 "$log"
    if [ $rc -ne 0 ]; then echo "[$tool] FAILED (rc $rc), see $log"; fail=$((fail + 1)); else echo "[$tool] clean"; fi
done
exit $fail
