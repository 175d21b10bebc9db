#!/bin/bash
# compute-sanitizer over the whole hot path (SURVEY.md §5 "race detection"): memcheck, racecheck, initcheck, synccheck on
# one small invocation of every entry point (tools/sanitize_workload.py).  GPU box only:
#
#     /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.txt 2>&1'
#
# The persistent kernel's dependency waits have a 2 s watchdog; under the sanitizer kernels run 10-100x slower, so the
# watchdog is raised (AGP_WAIT_TIMEOUT_MS).  All four tools are expected to be clean (profiles/r02_sanitize.txt).  What it
# took: racecheck does not model mbarriers (the chunk counter in shared memory is now written behind a CTA barrier),
# initcheck saw the pad bytes between lml[] and info[] of the result block (zeroed at allocation), synccheck wants every
# thread of a named barrier to execute the same BAR instruction (do_potf2: one out-of-line copy for all call sites).
set -u
cd "$(dirname "$0")/.."
export AGP_WAIT_TIMEOUT_MS=600000
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
for tool in ${SAN_TOOLS:-memcheck racecheck initcheck synccheck}; do
    echo "=== compute-sanitizer --tool $tool"
    timeout ${SAN_TIMEOUT:-420} $SAN --tool $tool --error-exitcode 9 --print-limit 5 python tools/sanitize_workload.py > /tmp/agp_san_$tool.txt 2>&1
    st=$?
    grep -v "^=========$\|Host Frame\|Saved host backtrace" /tmp/agp_san_$tool.txt | cut -c1-240 | awk 'NR <= 30 { print } { last[NR % 12] = $0 } END { if (NR > 30) { print "  ..."; for (i = NR - 11; i <= NR; ++i) if (i > 30) print last[i % 12] } }'
    echo "=== $tool exit status $st"
    if [ "$st" != 0 ]; then rc=1; fi
done
exit $rc
