cd $GRAFT_REPO_ROOT
python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "predict_gram_host_path" 2>&1 | tail -15
