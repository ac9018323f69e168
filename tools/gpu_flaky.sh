# Repeat the GPU suite to expose run-to-run flakiness (atomics order, clocks), then smoke + the default bench line.
set -x
cd $GRAFT_REPO_ROOT
T=${TAG:-r02h}
for i in 1 2 3 4; do
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${T}_pytest_$i.log 2>&1; echo "pytest run $i exit $?"; tail -2 gpurun_out/${T}_pytest_$i.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err; head -c 1500 gpurun_out/${T}_bench.json
