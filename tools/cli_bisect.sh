#!/bin/bash
# Developer tool: which round-2 feature makes the 2-rank CLI differ from the 1-rank CLI?  (2-GPU box)
cmp() { python - "$1" "$2" <<'PY'
import glob, sys, cv2, numpy as np
def frames(p):
    cap = cv2.VideoCapture(p); out = []
    while True:
        ok, img = cap.read()
        if not ok: break
        out.append(img)
    return np.stack(out)
a_dir, b_dir = sys.argv[1], sys.argv[2]
for kind in ("orig", "fifo"):
    a = frames(glob.glob(f"{a_dir}/outputs/*/clip1_{kind}_*.mp4")[0]); b = frames(glob.glob(f"{b_dir}/outputs/*/clip1_{kind}_*.mp4")[0])
    print("   ", kind, "identical" if np.array_equal(a, b) else f"DIFFERENT max|d|={np.abs(a.astype(int)-b.astype(int)).max()} frames differing={[int(i) for i in np.nonzero((a!=b).reshape(len(a),-1).any(1))[0]]}")
PY
}
run1() { rm -rf $1; python tools/make_tiny_checkpoint.py $1 > /dev/null; shift; }
mk() { d=$1; shift; rm -rf $d; python tools/make_tiny_checkpoint.py $d > /dev/null; for l in "$@"; do echo "$l" >> $d/tiny_edit.yaml; done; }
mk /tmp/c1; CUDA_VISIBLE_DEVICES=0 python infer_cogvideo_mp_fifo.py --config /tmp/c1/tiny_edit.yaml > /tmp/c1.log 2>&1 || tail -5 /tmp/c1.log
mk /tmp/c1b; CUDA_VISIBLE_DEVICES=0 python infer_cogvideo_mp_fifo.py --config /tmp/c1b/tiny_edit.yaml > /tmp/c1b.log 2>&1
echo "1 rank vs 1 rank again:"; cmp /tmp/c1 /tmp/c1b
two() { d=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) infer_cogvideo_mp_fifo.py --config $d/tiny_edit.yaml > $d.log 2>&1 || tail -5 $d.log; }
mk /tmp/c2a; two /tmp/c2a X=1; echo "2 ranks default (cfg-parallel base):"; cmp /tmp/c1 /tmp/c2a
mk /tmp/c2b "streaming_decode: false"; two /tmp/c2b X=1; echo "2 ranks, streaming_decode off:"; cmp /tmp/c1 /tmp/c2b
mk /tmp/c2c "ramp_sharding: false"; two /tmp/c2c X=1; echo "2 ranks, ramp_sharding off:"; cmp /tmp/c1 /tmp/c2c
mk /tmp/c2d "cfg_parallel: false"; two /tmp/c2d X=1; echo "2 ranks, cfg_parallel off:"; cmp /tmp/c1 /tmp/c2d
mk /tmp/c2e; two /tmp/c2e TG_VAE_FUSED_STATS=0; mk /tmp/c1e; CUDA_VISIBLE_DEVICES=0 TG_VAE_FUSED_STATS=0 python infer_cogvideo_mp_fifo.py --config /tmp/c1e/tiny_edit.yaml > /tmp/c1e.log 2>&1; echo "separate stats, 1 vs 2 ranks:"; cmp /tmp/c1e /tmp/c2e
mk /tmp/c2f "sequence_parallel: true"; two /tmp/c2f X=1; echo "2 ranks, sequence-parallel base:"; cmp /tmp/c1 /tmp/c2f
mk /tmp/c2g "cfg_parallel: false" "streaming_decode: false" "ramp_sharding: false"; two /tmp/c2g X=1; echo "2 ranks, everything off:"; cmp /tmp/c1 /tmp/c2g
