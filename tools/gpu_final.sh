# Final-tree confirmation: the GPU suite, smoke, the default bench line.
set -x
cd $GRAFT_REPO_ROOT
T=${TAG:-r02z}
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err; head -c 400 gpurun_out/${T}_bench.json
