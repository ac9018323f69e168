# A/B of two builds of the library on the training step: ab/libdmf_a.so (saved earlier) against the in-tree build.
cd $GRAFT_REPO_ROOT
cp dual-modal-fusion_b200/dmf/libdmf_b200.so /tmp/libdmf_b.so
for rep in 1 2; do
  for v in b a; do
    if [ $v = a ]; then cp ab/libdmf_a.so dual-modal-fusion_b200/dmf/libdmf_b200.so; else cp /tmp/libdmf_b.so dual-modal-fusion_b200/dmf/libdmf_b200.so; fi
    echo "== build $v rep $rep"; timeout 300 python tools/train_probe.py 512 50 2>&1 | tail -1 | cut -c100-260
  done
done
cp /tmp/libdmf_b.so dual-modal-fusion_b200/dmf/libdmf_b200.so
