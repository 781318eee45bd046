python tools/host_bound.py 2>&1 | tail -2
