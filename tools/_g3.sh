cd $GRAFT_REPO_ROOT
S=$SECONDS
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "suite seconds: $((SECONDS-S))"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
