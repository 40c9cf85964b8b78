timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python tools/exp_train_time.py 8 n 2>&1 | tail -3
