cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6; do python -m pytest tests/test_reference_pin.py -m gpu -q -k "bf16" 2>&1 | tail -1; done
