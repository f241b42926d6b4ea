cd $GRAFT_REPO_ROOT
for t in default 4 8 12 16; do
  if [ $t = default ]; then unset FRK_HOST_COPY_THREADS; else export FRK_HOST_COPY_THREADS=$t; fi
  python tools/predict_gram_host_probe.py
done
nproc; lscpu | grep -i "numa\|socket\|model name" | head
