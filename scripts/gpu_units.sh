#!/bin/bash
# timing + parity points for units above the one-stage limit: ~105 KB and ~150 KB units, 1 000-key batch,
# the default dispatch (tile ring, KIND mode) against the gather kernel; then the racecheck-free sanity of the defaults
mkdir -p gpurun_out
timeout 300 python scripts/sweep_tiles.py 600x15000 "BSG_PROBE_VARIANT=7" "SWEEP_PATH=gather" > gpurun_out/r02_units_105k.txt 2> gpurun_out/r02_units_105k.err; echo "105k rc=$?"; cat gpurun_out/r02_units_105k.txt
timeout 300 python scripts/sweep_tiles.py 450x21500 "BSG_PROBE_VARIANT=7" "SWEEP_PATH=gather" > gpurun_out/r02_units_150k.txt 2> gpurun_out/r02_units_150k.err; echo "150k rc=$?"; cat gpurun_out/r02_units_150k.txt
timeout 200 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r02_units_2b.txt 2> gpurun_out/r02_units_2b.err; echo "2b rc=$?"; cat gpurun_out/r02_units_2b.txt
timeout 200 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" > gpurun_out/r02_units_2a.txt 2> gpurun_out/r02_units_2a.err; echo "2a rc=$?"; cat gpurun_out/r02_units_2a.txt
