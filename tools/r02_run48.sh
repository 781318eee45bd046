timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "job_table_cache or keyswitch_and_rotations or mul_relin" 2>&1 | tail -3
