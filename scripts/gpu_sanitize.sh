# compute-sanitizer over one small render of every BASELINE scene (memcheck: out-of-bounds / misaligned accesses in
# global, shared and local memory; racecheck: shared-memory hazards around the TMA-staged scene)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import rtiow_rust_b200 as R
for name, nx, ny, ns, bvh in (("book1", 64, 48, 4, True), ("cornell", 32, 32, 4, False), ("final", 32, 32, 4, False), ("kitchen_sink", 32, 32, 4, True)):
    w, c = R.build_scene(name, nx, ny, use_bvh=bvh)
    for trav in (0, 2, 1):
        w.set_traversal(trav)
        img = R.par_cast(nx, ny, ns, c, w).rgb
    print(name, "ok", float(img.mean()))
    w.close()
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
