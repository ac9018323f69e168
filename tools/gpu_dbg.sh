cd $GRAFT_REPO_ROOT
for d in 0 1 2 3; do
  echo "DBG=$d"; DMF_DENSE_DBG=$d timeout 300 python tools/dense_probe.py 1000 1000 512 3 2>&1 | tail -1
done
