set -x
cd $GRAFT_REPO_ROOT
T=${TAG:-r02d}
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_train.py::test_data_parallel_gradient_parity_nccl -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python tools/datapath_probe.py > gpurun_out/${T}_datapath.jsonl 2> gpurun_out/${T}_datapath.err; echo "probe exit $?"; cat gpurun_out/${T}_datapath.jsonl; tail -3 gpurun_out/${T}_datapath.err
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${T}_bench.err; head -c 1200 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref exit $?"; cat gpurun_out/${T}_bench_reference.json | head -c 1500
timeout 300 python tools/datapath_probe.py --once > gpurun_out/${T}_once.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:'gather_tma|ihs_|argmax_confusion|paint|confusion_at|pan2ms' -c 60 -o gpurun_out/${T}_datapath python tools/datapath_probe.py --once > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/${T}_ncu.log
DENSE_ONCE=1 timeout 300 python tools/dense_probe.py 512 2101 512 1 > gpurun_out/${T}_dense_once.log 2>&1 && DENSE_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_map|conv_pool4|fuse_rowsum|head_dense' -c 10 -o gpurun_out/${T}_dense python tools/dense_probe.py 512 2101 512 1 > gpurun_out/${T}_ncu_dense.log 2>&1; echo "ncu dense exit $?"; tail -2 gpurun_out/${T}_ncu_dense.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
