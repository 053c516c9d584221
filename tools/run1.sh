python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_eval.py 4 20 1e7 chebyshev
python tools/quick_eval.py 4 20 1e7
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
