#!/bin/bash
# compute-sanitizer over smoke() and the ragged / edge-case parity tests (race / memory checking is a
# net-new deliverable).  One log per tool under gpurun_out/, the summaries go to profiles/sanitizer_rNN.md.
#   bash tools/sanitize.sh [memcheck racecheck initcheck synccheck]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-"memcheck racecheck initcheck synccheck"}
TESTS='tests/test_gpu_parity.py -k "edge_lengths or trim_disabled or crossfade_disabled or ragged_lengths or compact or joined_items or silence_and_constant or cosine"'
for tool in $TOOLS; do
  log=gpurun_out/sanitizer_r02_${tool}.log
  echo "=== compute-sanitizer --tool $tool: smoke()" > $log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -c "import __graft_entry__ as g; g.smoke()" >> $log 2>&1
  echo "=== compute-sanitizer --tool $tool: pytest $TESTS" >> $log
  eval timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -m pytest $TESTS -m gpu -q -x -p no:cacheprovider >> $log 2>&1
  echo "--- $tool: $(grep -c 'ERROR SUMMARY' $log) runs; summaries:"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|smoke ok" $log
done
