#!/bin/bash
# call 64: launch-floor probe (parameter size, set-up, PDL) for the grouped engine's per-launch cost
mkdir -p gpurun_out
timeout 120 profiles/scripts/probes/launch_floor_probe.bin > gpurun_out/r2c64_launch_floor.txt 2>&1
cat gpurun_out/r2c64_launch_floor.txt
