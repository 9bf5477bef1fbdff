for m in 0 64 128 192; do echo "=== mode $m"; HB200_BAND_DBGMODE=$m timeout 120 python tools/band_timing.py 2>/dev/null | grep -v "^ \|^chain\|^step"; done
