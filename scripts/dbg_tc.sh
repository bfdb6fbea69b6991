for d in 7 23 0 16; do echo "DBG=$d"; MI_TC_DBG=$d timeout 60 python scripts/sweep_tc.py 2>/dev/null | sed -n 4,5p; done
