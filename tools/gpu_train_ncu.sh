cd $GRAFT_REPO_ROOT
T=${TAG:-r02t}
timeout 300 python tools/train_probe.py 512 20 > gpurun_out/${T}_train.json 2>&1; echo "probe exit $?"; tail -1 gpurun_out/${T}_train.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_bwd|bn_apply|pan1_' --launch-skip 150 -c 24 -o gpurun_out/${T}_train python tools/train_probe.py 512 6 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/${T}_ncu.log
