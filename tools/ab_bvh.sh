python -m pytest tests/test_traversal_parity.py tests/test_render_parity.py -m gpu -q -x 2>&1 | tail -3
for w in materials cornell terrain; do
  steps=32; [ $w = terrain ] && steps=8
  for bvh in lbvh ploc; do
    BPT_BVH=$bvh python bench.py --steps $steps --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w $bvh', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms  extend', round(r['share_of_step']['extend']*d['ms_per_step'],3), 'shadow', round(r['share_of_step']['shadow']*d['ms_per_step'],3), 'build_ms', round(d['bvh']['build_ms'],2), 'nodes', d['bvh']['nodes'])"
    BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_stats.so BPT_BVH=$bvh python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   stats $w $bvh nodes/ray', round(d['extend_node_visits_per_ray'],1), 'tris/ray', round(d['extend_triangle_tests_per_ray'],1))"
  done
done
