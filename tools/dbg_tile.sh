for d in 0 1 3 5; do SAME_B200_TILE_DBG=$d python tools/time_candidates.py 2500 4 2>&1 | grep -E "k_knn|pairs"; done
