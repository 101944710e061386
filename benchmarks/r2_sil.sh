timeout 900 python -m pytest tests/test_silhouette.py -m gpu -x -q --timeout 600 -s 2>&1 | grep -v "^$" | tail -25
