"""Self-check of the tile-sorted convolution path (GCDLSS_TILE_SORT) on the GPU it will run on.

bench.py runs this in a SUBPROCESS before its own measurement (a failure, a device fault or a hang here cannot touch the
bench's CUDA context) and switches tile sorting on only if every check passes and the sorted path is measurably faster:

  1. gcd_kmap_tile_sort == a stable torch sort of the presence-mask keys (3x3x3 and 2x2x2 tables), bit exact;
  2. forward and dgrad convolutions through sorted tables against an fp64 re-computation, within the stated bf16 bound;
  3. one training step of the bench's model in both modes: loss and logits agree;
  4. forward + backward time on ready maps in both modes (the bench builds maps and sort on a side stream, one batch
     ahead); the map-build times are reported next to it.

Prints one JSON line: {"ok": bool, "reason": str, ...}.  Exit code 0 whenever the line was printed."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _paths  # noqa: E402,F401

import torch  # noqa: E402

TOL_BF16 = 3e-2          # tests/gpu_util.py


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def presence_bits(kv):
    """Bit of each kernel offset in the sort key (csrc/tilesort.cuh): 3x3x3 offsets ranked by (number of non-zero
    components, k) -- centre lowest, corners on top; 2x2x2: bit k."""
    if kv != 27:
        return torch.arange(kv)
    k = torch.arange(27)
    cls = ((k % 3) != 1).long() + (((k // 3) % 3) != 1).long() + ((k // 9) != 1).long()
    order = torch.argsort(cls * 27 + k)
    bits = torch.empty(27, dtype=torch.long)
    bits[order] = torch.arange(27)
    return bits


def main():
    device = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    scans = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    kind = sys.argv[3] if len(sys.argv) > 3 else "kitti"
    n_classes = int(sys.argv[4]) if len(sys.argv) > 4 else 17
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    import gcdlss_b200
    import MinkowskiEngine as ME
    import bench
    from gcdlss_b200 import ops, synth
    from gcdlss_b200.steps import point_cross_entropy
    from models.multiheadminkunet import MinkUNetBase

    out = {"ok": False, "reason": "", "device": device}
    gcdlss_b200.set_math_mode("bf16")
    host = bench.make_host_batches(kind, scans, None, n_classes, 0, 1)
    bc, f, labels = bench.quantize_batch_on_gpu(host[0], synth.voxel_size(kind), dev)

    # ---- 1. the sort itself
    gcdlss_b200.set_tile_sort(False)
    mgr = ME.SparseTensor(features=f, coordinates=bc).coordinate_manager
    for key in ((1, 3, 1, False), (2, 3, 1, False), (1, 2, 2, False), (2, 2, 2, True)):
        km = mgr.kernel_map(*key)
        nbr = km.nbr
        kv = nbr.shape[0]
        bits = presence_bits(kv).to(dev)
        keys = ((nbr >= 0).long() << bits[:, None]).sum(0)
        ref_rows = torch.argsort(keys, stable=True)
        got, rows = ops.kmap_tile_sort(nbr)
        if not (torch.equal(rows.long(), ref_rows) and torch.equal(got, nbr[:, ref_rows])):
            out["reason"] = f"sorted table of map {key} differs from the torch reference"
            return out

    # ---- 2. convolutions through sorted tables vs fp64
    g = torch.Generator(device=dev).manual_seed(1)
    worst = 0.0
    for key, cin, cout in (((1, 3, 1, False), 96, 96), ((2, 3, 1, False), 32, 64), ((2, 2, 2, True), 64, 32), ((1, 2, 2, False), 32, 32)):
        km = mgr.kernel_map(*key)
        table, rows = ops.kmap_tile_sort(km.nbr)
        kv = km.nbr.shape[0]
        x = torch.randn(km.n_in, cin, device=dev, generator=g).to(torch.bfloat16)
        w = torch.randn(kv, cin, cout, device=dev, generator=g) * 0.05
        y = ops.conv_forward(x, table, w, km.n_out, out_dtype=torch.bfloat16, math_mode=1, w_packed=ops.pack_weights(w, False, False), out_rows=rows)
        ref = torch.zeros(km.n_out, cout, dtype=torch.float64, device=dev)
        for k in range(kv):
            idx = km.nbr[k].long()
            o = torch.nonzero(idx >= 0).reshape(-1)
            ref.index_add_(0, o, x.double()[idx[o]] @ w[k].double())
        worst = max(worst, rel(y, ref))
    out["conv_rel_err"] = worst
    if not worst < TOL_BF16:
        out["reason"] = f"convolution through a sorted table off by {worst:.3e} (bound {TOL_BF16})"
        return out

    # ---- 3 + 4. one model, both modes: agreement and step time
    torch.manual_seed(1234)
    model = MinkUNetBase(num_classes=n_classes).to(dev).train()

    def build():
        st = ME.SparseTensor(features=f, coordinates=bc)
        st.coordinate_manager.prebuild_unet(5, 5, with_pairs=True)
        return st

    def step(st):
        model.zero_grad(set_to_none=True)
        logits = model(st)["logits"]
        loss = point_cross_entropy(logits, labels)
        loss.backward()
        return logits.detach().float(), float(loss)

    def timed(fn, reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # The bench builds the maps (and the sort) of batch i + 1 on a side stream while batch i trains, so the step it measures is
    # the forward + backward on ready maps: that is what is compared here; the cost of the build is reported next to it.
    res = {}
    for mode in (False, True, False, True):          # interleaved, the second pass of each mode is the one that counts
        gcdlss_b200.set_tile_sort(mode)
        st = build()
        for _ in range(3):
            logits, loss = step(st)
        ms_step = timed(lambda: step(st), 5)
        ms_build = timed(build, 3)
        res[mode] = (logits, loss, ms_step, ms_build)
    gcdlss_b200.set_tile_sort(False)
    (l0, loss0, ms0, mb0), (l1, loss1, ms1, mb1) = res[False], res[True]
    out.update(ms_scan_order=ms0, ms_sorted=ms1, ms_build_scan_order=mb0, ms_build_sorted=mb1, loss_scan_order=loss0, loss_sorted=loss1,
               logits_rel_diff=rel(l1, l0))
    if not (abs(loss0 - loss1) < 1e-2 and out["logits_rel_diff"] < 5e-2 and loss1 == loss1):
        out["reason"] = "training step disagrees between the sorted and the scan-order path"
        return out
    if not ms1 < 0.97 * ms0:
        out["reason"] = f"sorted path not faster ({ms1:.2f} ms vs {ms0:.2f} ms per step)"
        return out
    out["ok"] = True
    out["reason"] = f"checks passed; forward + backward {ms0:.2f} -> {ms1:.2f} ms, map build {mb0:.2f} -> {mb1:.2f} ms (side stream in the bench)"
    return out


if __name__ == "__main__":
    t0 = time.time()
    try:
        result = main()
    except Exception as e:  # noqa: BLE001 -- any failure means "leave it off"
        result = {"ok": False, "reason": f"{type(e).__name__}: {e}"}
    result["seconds"] = round(time.time() - t0, 1)
    print(json.dumps(result), flush=True)
