# Round-2 evidence of the dense path after the sub-position sharing: tests, smoke, bench, one ncu --set full capture, the launch list.
set -x
cd $GRAFT_REPO_ROOT
T=${TAG:-r02k}
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err; head -c 1800 gpurun_out/${T}_bench.json
DENSE_ONCE=1 timeout 300 python tools/dense_probe.py 512 2101 512 1 > gpurun_out/${T}_dense_once.log 2>&1 && DENSE_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_map|conv_pool4|fuse_rowsum|head_dense' -c 10 -o gpurun_out/${T}_dense python tools/dense_probe.py 512 2101 512 1 > gpurun_out/${T}_ncu_dense.log 2>&1; echo "ncu dense exit $?"; tail -2 gpurun_out/${T}_ncu_dense.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
