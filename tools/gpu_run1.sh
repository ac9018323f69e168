set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x --deselect tests/test_gpu_train.py::test_data_parallel_gradient_parity_nccl > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -30 gpurun_out/r02a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02a_smoke.log
timeout 600 python tools/datapath_probe.py > gpurun_out/r02a_datapath.jsonl 2> gpurun_out/r02a_datapath.err; echo "probe exit $?"; cat gpurun_out/r02a_datapath.jsonl; tail -5 gpurun_out/r02a_datapath.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"; tail -5 gpurun_out/r02a_bench.err; head -c 3000 gpurun_out/r02a_bench.json
timeout 300 python tools/datapath_probe.py --once > gpurun_out/r02a_once.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:'gather_tma|ihs_|argmax_confusion|paint|confusion_at|pan2ms' -c 40 -o gpurun_out/r02a_datapath python tools/datapath_probe.py --once > gpurun_out/r02a_ncu.log 2>&1; echo "ncu exit $?"; tail -5 gpurun_out/r02a_ncu.log
