python scripts/e2e_probe.py 2>&1 | tail -20
