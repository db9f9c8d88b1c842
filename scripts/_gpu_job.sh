mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python scripts/profile_generic_D.py --dmax 16 2>/dev/null | tee gpurun_out/r3h_generic_D16.json
timeout 300 python scripts/profile_generic_D.py --dmax 8 2>/dev/null | tee gpurun_out/r3h_generic_D8.json
