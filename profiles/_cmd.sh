python profiles/windows_scan.py 200000 102 2>&1 | tail -2 | cut -c1-700
