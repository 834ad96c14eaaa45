#!/bin/bash
# Diagnosis of k_linearize_landmark4 (needs a library built with CUBA_JH4_DEBUG=1 python cuda-bundle-adjustment_b200/build.py):
#   CUBA_JH4_DBG bit 0 skips the arithmetic, bit 1 the Hpl staging + bulk store, bit 2 the per-landmark reduction,
#   bit 3 adds clock64 phase counters and globaltimer marks (printed to stderr on the fourth launch).
for d in 0 1 2 4 6 7 8; do
	echo "DBG $d"
	CUBA_JH4_DBG=$d timeout 100 python tools/jh_variants.py --variants 0 "${1:-kitti00_shaped}" 2>&1 | tail -11
done
