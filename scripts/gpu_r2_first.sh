#!/bin/bash
# round 2, first GPU call: Go probe, parity of the tile-ring kernel, full GPU suite, kernel sweeps
mkdir -p gpurun_out
{ echo "== go probe =="; which go; go version; ls -d /usr/local/go ~/go /root/go/pkg/mod 2>&1; env | grep -i "^GO"; echo "nproc $(nproc)"; nvidia-smi -L; } > gpurun_out/r2_goprobe.txt 2>&1
cat gpurun_out/r2_goprobe.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles or beyond or boundaries or small_corpus" > gpurun_out/r2_pytest_tiles.log 2>&1; echo "pytest(tiles) rc=$?"; tail -15 gpurun_out/r2_pytest_tiles.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -5 gpurun_out/r2_pytest_all.log
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=3" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=1" "BSG_TILES_SHAPE=2" "BSG_TILES_SHAPE=3" "BSG_TILES_SHAPE=4" "BSG_TILES_SHAPE=5" "BSG_TILES_SHAPE=6" "BSG_TILES_SHAPE=0 BSG_TILE_MODE=1" "BSG_TILES_SHAPE=0 BSG_PROBE_PDL=0" > gpurun_out/r2_sweep_2b.txt 2> gpurun_out/r2_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2_sweep_2b.txt; tail -3 gpurun_out/r2_sweep_2b.err
timeout 900 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=3" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=1" "BSG_TILES_SHAPE=2" "BSG_TILES_SHAPE=3" "BSG_TILES_SHAPE=4" "BSG_TILES_SHAPE=5" "BSG_TILES_SHAPE=6" "BSG_TILES_SHAPE=0 BSG_TILE_BYTES=16384" "BSG_TILES_SHAPE=0 BSG_TILE_BYTES=60000" "BSG_TILES_SHAPE=0 BSG_TILE_UNITS=2" "BSG_TILES_SHAPE=0 BSG_TILE_UNITS=1" "BSG_TILES_SHAPE=1 BSG_TILE_BYTES=16384" "BSG_TILES_SHAPE=0 BSG_PROBE_PDL=0" > gpurun_out/r2_sweep_2a.txt 2> gpurun_out/r2_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2_sweep_2a.txt; tail -3 gpurun_out/r2_sweep_2a.err
