timeout 600 python -m pytest tests -q -m gpu -x -k "detection_loss" 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
