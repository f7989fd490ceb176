# The CLI on 1 rank vs 2 ranks (CFG-parallel base clip + window-parallel FIFO + clip-parallel decode) vs 2 ranks with the
# sequence-parallel base clip (`sequence_parallel: true`): identical videos.
set -e
rm -rf /tmp/ckA /tmp/ckB /tmp/ckC
python tools/make_tiny_checkpoint.py /tmp/ckA > /dev/null
python tools/make_tiny_checkpoint.py /tmp/ckB > /dev/null
python tools/make_tiny_checkpoint.py /tmp/ckC > /dev/null
echo "sequence_parallel: true" >> /tmp/ckC/tiny_edit.yaml
CUDA_VISIBLE_DEVICES=0 python infer_cogvideo_mp_fifo.py --config /tmp/ckA/tiny_edit.yaml > /tmp/cliA.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 infer_cogvideo_mp_fifo.py --config /tmp/ckB/tiny_edit.yaml > /tmp/cliB.log 2>&1 || { tail -20 /tmp/cliB.log; exit 1; }
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 infer_cogvideo_mp_fifo.py --config /tmp/ckC/tiny_edit.yaml > /tmp/cliC.log 2>&1 || { tail -20 /tmp/cliC.log; exit 1; }
python - <<'PY'
import glob, cv2, numpy as np
def frames(p):
    cap = cv2.VideoCapture(p); out = []
    while True:
        ok, img = cap.read()
        if not ok: break
        out.append(img)
    return np.stack(out)
for kind in ("orig", "fifo"):
    a = frames(glob.glob(f"/tmp/ckA/outputs/*/clip1_{kind}_*.mp4")[0]); b = frames(glob.glob(f"/tmp/ckB/outputs/*/clip1_{kind}_*.mp4")[0])
    print(kind, a.shape, "identical" if np.array_equal(a, b) else f"DIFFERENT max|d|={np.abs(a.astype(int)-b.astype(int)).max()}")
    c = frames(glob.glob(f"/tmp/ckC/outputs/*/clip1_{kind}_*.mp4")[0])
    print(kind, "sequence-parallel base clip:", "identical" if np.array_equal(a, c) else f"DIFFERENT max|d|={np.abs(a.astype(int)-c.astype(int)).max()}")
PY
