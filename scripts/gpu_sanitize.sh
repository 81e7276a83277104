# compute-sanitizer over one small render of every BASELINE scene (memcheck: out-of-bounds / misaligned accesses in
# global, shared and local memory; racecheck: shared-memory hazards around the TMA-staged scene)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import rtiow_rust_b200 as R
import torch
from rtiow_rust_b200 import dist as rdist
QUICK = bool(os.environ.get("SAN_QUICK"))
for name, nx, ny, ns, bvh in (("book1", 64, 48, 4, True), ("final", 32, 32, 4, False)) if QUICK else (("book1", 64, 48, 4, True), ("cornell", 32, 32, 4, False), ("final", 32, 32, 4, False), ("kitchen_sink", 32, 32, 4, True),
                              ("cornell_smoke", 32, 32, 4, False), ("simple_light", 32, 32, 2, True)):
    w, c = R.build_scene(name, nx, ny, use_bvh=bvh)
    for trav in (0, 2, 1):
        w.set_traversal(trav)
        img = R.par_cast(nx, ny, ns, c, w).rgb          # consecutive renders alternate between the two pipeline slots
    q = R.par_cast_ppm(nx, ny, ns, c, w)
    print(name, "ok", float(img.mean()), int(q.sum()))
    w.close()
# a frame with enough tiles for the learnt unit order (>= 1024): renders 3 and 4 take their tiles from the sorted strip table
# that every CTA copies into shared memory behind the scene
nx, ny, ns = 320, 208, 2
w, c = R.build_scene("book1", nx, ny)
imgs = [R.par_cast(nx, ny, ns, c, w).rgb for _ in range(4)]
print("learnt order ok", all((imgs[0] == i).all() for i in imgs))
w.close()
if os.environ.get("SAN_QUICK"):
    sys.exit(0)
# the multi-GPU exchange with both ranks on this GPU: fold stores into two frames, flag hand-shake
nx, ny, ns = 64, 48, 3
worlds = [R.build_scene("book1", nx, ny) for _ in range(2)]
frames = [rdist.PeerFrame(worlds[r][0].lib, nx, ny, 0, r, 2, connect=False) for r in range(2)]
for f in frames:
    f.connect([g.handle for g in frames])
streams = [torch.cuda.Stream() for _ in range(2)]
for rep in range(3):
    for r in range(2):
        frames[r].render(nx, ny, ns, worlds[r][1], worlds[r][0], stream=streams[r])
    for st in streams:
        st.synchronize()
print("peers ok", float(frames[0].frame.mean()), bool((frames[0].frame == frames[1].frame).all()))
PY
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
