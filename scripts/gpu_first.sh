set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
python -m pytest tests -x -q -m gpu 2>&1 | tail -30
python - <<'PY'
import sys,time; sys.path.insert(0,'.')
import rtiow_rust_b200 as R
for name,nx,ny,ns,bvh in (("book1",1200,800,50,True),("cornell",800,800,100,False),("final",800,800,100,False)):
    w,c=R.build_scene(name,nx,ny,use_bvh=bvh)
    for thr,cps in ((256,0),(128,0),(512,0),(256,1)):
        w.set_tuning(cta_threads=thr, ctas_per_sm=cps)
        for rep in range(2):
            t=time.time(); R.par_cast(nx,ny,ns,c,w); dt=time.time()-t
        st=w.stats()
        print(f"{name} {nx}x{ny}x{ns} thr={thr} cps={cps}: trace {st['trace_ms']:.2f} ms fold {st['reduce_ms']:.2f} ms -> {st['samples']/st['trace_ms']/1e3:.1f} Msamples/s (kernel), e2e {dt*1e3:.1f} ms, grid {st['grid']} regs {st['regs_per_thread']} smem {st['dyn_smem_bytes']} segs/sample {st['segments']/st['samples']:.3f}")
PY
