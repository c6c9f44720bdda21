timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -1
