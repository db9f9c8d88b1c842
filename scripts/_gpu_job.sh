mkdir -p gpurun_out
timeout 900 python scripts/run_config1_maxcut.py 1000 --measure 2>/dev/null | tee gpurun_out/r3g_config1_full.json | cut -c100-900
