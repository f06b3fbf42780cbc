#!/bin/bash
timeout -k 5 900 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=900 -k "batched_equals and 1-27" 2>&1 | grep -E "error detected|Device Frame|ERROR SUMMARY" | head -20
echo ----
timeout -k 5 900 compute-sanitizer --tool synccheck --print-limit 4 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=900 -k "every_fast_variant" 2>&1 | grep -E "error detected|Device Frame|ERROR SUMMARY" | head -24
