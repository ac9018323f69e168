set -x
cd $GRAFT_REPO_ROOT
T=${TAG:-r02i}
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity_fitted.py -x -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${T}_pytest.log
timeout 300 python tools/dense_probe.py 1000 1000 512 3 > gpurun_out/${T}_probe_c2.json 2>gpurun_out/${T}_probe_c2.err; echo "probe exit $?"; cat gpurun_out/${T}_probe_c2.json; tail -3 gpurun_out/${T}_probe_c2.err
DMF_DENSE_SHARE=0 timeout 300 python tools/dense_probe.py 1000 1000 512 3 > gpurun_out/${T}_probe_c2_noshare.json 2>&1; cat gpurun_out/${T}_probe_c2_noshare.json
timeout 300 python tools/dense_probe.py 2001 2101 512 3 > gpurun_out/${T}_probe_c3.json 2>&1; cat gpurun_out/${T}_probe_c3.json
