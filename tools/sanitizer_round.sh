#!/bin/bash
# compute-sanitizer memcheck over the small render / AOV / texture / traversal / incremental-update tests on one B200
# (the full-size tests are left out: the tool slows kernels down by one to two orders of magnitude). The log goes to
# gpurun_out/<tag>_sanitizer_memcheck.log; copy it to profiles/. The kernels run as plain stream launches (BPT_GRAPH=0): under the
# tool a graph with a conditional WHILE node ends in cudaErrorIllegalAddress without any memcheck record (no kernel, no address),
# also for graphs whose kernels are clean as stream launches.
R=${1:-r02}; O=gpurun_out; mkdir -p $O
K="cornell_box_matches or progressive_accumulation or vertex_tints or environment_map_importance or terrain_small or aov_backends or transmissive_shading_model or textured_materials or interval_and or deep_hierarchy or transform_edit or every_hierarchy or degenerate_geometry or environment_cdf_next_event or hit_sorting"
BPT_GRAPH=${BPT_GRAPH:-0} timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 --log-file $O/${R}_sanitizer_memcheck.log \
  python -m pytest tests/test_render_parity.py tests/test_textures.py tests/test_traversal_parity.py tests/test_incremental_updates.py -m gpu -q -x -k "$K" > $O/${R}_sanitizer_pytest.log 2>&1
echo "sanitizer run rc=$?"; tail -3 $O/${R}_sanitizer_pytest.log; grep -E "ERROR SUMMARY|Invalid|Error" $O/${R}_sanitizer_memcheck.log | head -10
