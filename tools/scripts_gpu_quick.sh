for v in 1 0; do echo "ATT_N4=$v"; PSIF_ATT_N4=$v python tools/eloc_error_stats.py LiH 256; PSIF_ATT_N4=$v python tools/eloc_error_stats.py Be 256; done
