for v in 0 8 7; do CUBA_JH4_NOBIG=1 timeout 100 python tools/jh_variants.py --variants $v kitti00_shaped 2>&1 | tail -1; done
for d in 8 7; do echo "DBG $d"; CUBA_JH4_NOBIG=1 CUBA_JH4_DBG=$d timeout 100 python tools/jh_variants.py --variants 0 kitti00_shaped 2>&1 | tail -11; done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jh_landmark or stage_parity or fixed_vertices" 2>&1 | tail -3
