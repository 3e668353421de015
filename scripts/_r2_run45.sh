python tests/_sweep.py 2>&1 | tail -12
