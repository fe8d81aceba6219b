python scripts/e2e_probe.py 2>&1 | grep -E "submit|host|two" 
