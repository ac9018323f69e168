# A/B of two builds of the library on one box: ab/libdmf_a.so (saved earlier) against the in-tree build.  ARGS = dense_probe arguments.
cd $GRAFT_REPO_ROOT
ARGS=${ARGS:-"1000 1000 512 5"}
cp dual-modal-fusion_b200/dmf/libdmf_b200.so /tmp/libdmf_b.so
for rep in ${REPS:-1 2}; do
  for v in b a; do
    if [ $v = a ]; then cp ab/libdmf_a.so dual-modal-fusion_b200/dmf/libdmf_b200.so; else cp /tmp/libdmf_b.so dual-modal-fusion_b200/dmf/libdmf_b200.so; fi
    echo "== build $v rep $rep"; timeout 300 python tools/dense_probe.py $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['dense_ms'],3), d['dense_stage_ms'])"
  done
done
cp /tmp/libdmf_b.so dual-modal-fusion_b200/dmf/libdmf_b200.so
