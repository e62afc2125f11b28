#!/bin/bash
# sanitizer over the round-2 kernels; register / occupancy variants of the cfg 5 kernel
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_detector.py tests/test_gpu_coatings.py tests/test_gpu_models.py "tests/test_gpu_grid.py::test_chromatic_generated_rays_match_oracle" "tests/test_gpu_grid.py::test_chromatic_fused_image_matches_oracle_and_the_specialised_kernel" "tests/test_gpu_grid.py::test_image_with_a_chromatic_stop_solution" "tests/test_gpu_grid.py::test_fused_grid_image_matches_oracle" -m gpu -q -x > gpurun_out/r02_sanitizer_memcheck_tests.txt 2>&1
echo "exit code $?" >> gpurun_out/r02_sanitizer_memcheck_tests.txt
tail -6 gpurun_out/r02_sanitizer_memcheck_tests.txt | cut -c 1-200
python bench.py --only-strong --strong cfg5 > gpurun_out/cfg5_default.json 2>/dev/null
OPTK_TRACE_HEAVY=0 OPTK_JIT_CACHE= python bench.py --only-strong --strong cfg5 > gpurun_out/cfg5_noheavy.json 2>/dev/null
OPTK_TRACE_HEAVY=1 OPTK_JIT_MINB=2 OPTK_JIT_CACHE= python bench.py --only-strong --strong cfg3 > gpurun_out/cfg3_heavy.json 2>/dev/null
python tools/gpu/summarize_strong.py gpurun_out/cfg5_default.json gpurun_out/cfg5_noheavy.json gpurun_out/cfg3_heavy.json
