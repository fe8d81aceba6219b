python -m pytest tests -m gpu -x -q -k "find or multi_gpu or pipelined" 2>&1 | tail -4
python scripts/e2e_probe2.py 2>&1 | tail -10
