echo '$ compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -k "not staged_upload and not fullsize and not full_size" 2>&1 | tail -4
echo '$ compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_span.py tests/test_gpu_volume.py -k "not random ..."'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_span.py tests/test_gpu_volume.py -q -m gpu -k "not random and not benchmark_resolution and not staged_upload and not zero_ness and not alongside" 2>&1 | tail -4
