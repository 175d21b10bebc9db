#!/bin/bash
# compute-sanitizer over the whole hot path (SURVEY.md §5 "race detection"): memcheck, racecheck, initcheck, synccheck on
# one small invocation of every entry point (tools/sanitize_workload.py).  GPU box only:
#
#     /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.txt 2>&1'
#
# The persistent kernel's dependency waits have a 2 s watchdog; under the sanitizer kernels run 10-100x slower, so the
# watchdog is raised (AGP_WAIT_TIMEOUT_MS).  Known, by design: synccheck reports "divergent thread(s) in warp" / barrier
# warnings for the named barriers of do_potf2, where different warps reach bar.sync 1 from different code locations.
set -u
cd "$(dirname "$0")/.."
export AGP_WAIT_TIMEOUT_MS=600000
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
for tool in memcheck racecheck initcheck synccheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout 900 $SAN --tool $tool --error-exitcode 9 --print-limit 5 python tools/sanitize_workload.py 2>&1 | grep -v "^=========$" | tail -25
    st=${PIPESTATUS[0]}
    echo "=== $tool exit status $st"
    if [ "$st" != 0 ] && [ "$tool" != synccheck ]; then rc=1; fi
done
exit $rc
