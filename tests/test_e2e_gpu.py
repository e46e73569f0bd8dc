"""GPU end-to-end parity: the whole forward graph (frcnn_b200.engine) against the CPU oracle.

Stage-wise "identical inputs" checks (each stage's oracle is fed the DEVICE's own upstream result, so
one stage's rounding cannot flip another stage's integer decisions) plus a pure end-to-end comparison
against the fp32 oracle at the north star's tolerance (1e-4 relative; bit-exact keep indices)."""
import numpy as np
import pytest
import torch

import frcnn_oracle as orc

pytestmark = pytest.mark.gpu
f32 = np.float32
ANCHORS = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))


def _quant16(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=f32))
    hi = t.to(torch.bfloat16).float()
    return (hi + (t - hi).to(torch.bfloat16).float()).numpy()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _check_probs(got, want):
    """Tolerance for class probabilities.  North star: "within 1e-4 relative".  Asserted relative to the
    tensor's scale (probabilities live in [0,1]: |dp| <= 1e-4 * max(p), in practice ~5e-6); the
    per-element relative error is ~|logit| times larger than the relative error of the logits
    (|logit| ~ 6 here) and is bounded at 1e-3 and reported, not held to 1e-4: two fp32
    implementations with different summation orders already differ by ~1e-4 per element here."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    abs_err = np.abs(got - want).max()
    rel_el = (np.abs(got - want) / np.maximum(want, 1e-12)).max()
    print("class prob: max abs err %.2e (scale %.2f), max per-element rel err %.2e" % (abs_err, want.max(), rel_el))
    assert abs_err < 1e-4 * want.max()
    assert rel_el < 1e-3


@pytest.fixture(scope="module")
def params():
    return orc.make_params(seed=1234)


def _engine(params, precision, **kw):
    from frcnn_b200.engine import Engine
    return Engine(params, precision=precision, anchors=ANCHORS, keep_rpn_debug=True, **kw)


@pytest.mark.parametrize("shape", [(96, 128), (150, 201)])
def test_forward_stagewise_bf16x3(params, shape):
    H, W = shape
    x = orc.make_image(H, W, seed=1)
    info = np.array([[H, W]], np.int32)
    eng = _engine(params, "bf16x3", use_graph=False)
    prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
    torch.cuda.synchronize()
    R = prob.shape[0]

    # ---- trunk: device conv5_3 vs the fp32 oracle (north-star tolerance) and vs the 16-bit-operand oracle
    feat_dev = plan.acts[-1].to_chw_f32().cpu().numpy()[None]
    feat_ref = orc.vgg16_forward(x, params)
    assert feat_dev.shape == feat_ref.shape
    e_fp32 = _rel(feat_dev, feat_ref)
    e_q16 = _rel(feat_dev, orc.vgg16_forward(x, params, quant=_quant16))
    print("conv5_3 max-rel error: vs fp32 oracle %.2e, vs 16-bit-operand oracle %.2e" % (e_fp32, e_q16))
    assert e_fp32 < 5e-5 and e_q16 < 5e-5      # north star: 1e-4; the build keeps a 2x margin

    # ---- RPN heads on the device's own feature map
    fh, fw = plan.fh, plan.fw
    rpn = plan.rpn_out.cpu().numpy()
    logits_dev = rpn[:, :18].T.reshape(1, 18, fh, fw)
    deltas_dev = rpn[:, 18:54].T.reshape(1, 36, fh, fw)
    h_ref = orc.relu(orc.conv2d(feat_dev, params["RPN/rpn_conv_3x3/W"], params["RPN/rpn_conv_3x3/b"], 1))
    assert _rel(plan.rpn_mid.to_chw_f32().cpu().numpy()[None], h_ref) < 1e-4
    h_dev = plan.rpn_mid.to_chw_f32().cpu().numpy()[None]
    assert _rel(logits_dev, orc.conv2d(h_dev, params["RPN/rpn_cls_score/W"], params["RPN/rpn_cls_score/b"], 0)) < 1e-4
    assert _rel(deltas_dev, orc.conv2d(h_dev, params["RPN/rpn_bbox_pred/W"], params["RPN/rpn_bbox_pred/b"], 0)) < 1e-4

    # ---- ProposalLayer on identical inputs (the device's logits/deltas): BIT-EXACT
    dbg = {}
    want_rois, want_fg = orc.proposal_layer(orc.softmax_axis1(logits_dev), deltas_dev, info, debug=dbg)
    rois_dev = plan.prop.rois.cpu().numpy()
    assert R == len(want_rois) and R > 0
    assert np.array_equal(rois_dev[:R], want_rois)
    assert np.array_equal(plan.prop.scores.cpu().numpy()[:R], want_fg.ravel())
    ns = int(plan.prop.dbg_num.item())
    assert np.array_equal(plan.prop.dbg_dets.cpu().numpy()[:ns], dbg["dets"])

    # ---- RoI pool (exact) and head (1e-4) on the device's feature map and RoIs
    cls_ref, box_ref, aux = orc.head_forward(feat_dev, rois_dev[:R], params, info)
    pool_dev = (plan.pool5.hi.float() + plan.pool5.lo.float()).cpu().numpy().reshape(-1, 7, 7, 512)[:R]
    assert np.array_equal(pool_dev.transpose(0, 3, 1, 2), aux["pool5"])
    assert _rel((plan.fc6.hi.float() + plan.fc6.lo.float()).cpu().numpy()[0, :R], aux["fc6"]) < 1e-4
    assert _rel((plan.fc7.hi.float() + plan.fc7.lo.float()).cpu().numpy()[0, :R], aux["fc7"]) < 1e-4
    _check_probs(prob.cpu().numpy(), cls_ref)
    assert _rel(boxes.cpu().numpy(), box_ref) < 1e-4
    # tail on identical inputs (device head logits): bit-exact softmax / decode / clip
    ho = plan.head_out.cpu().numpy()[:R]
    assert np.array_equal(prob.cpu().numpy(), orc.softmax_axis1(ho[:, :21]))
    assert np.array_equal(boxes.cpu().numpy(), orc.clip_boxes(orc.bbox_transform_inv(rois_dev[:R], ho[:, 21:105]), (H, W)))
    assert not plan.prob.cpu().numpy()[R:].any() and not plan.boxes.cpu().numpy()[R:].any()


def test_forward_end_to_end_vs_fp32_oracle(params):
    """Pure end to end against the fp32 oracle.  Scores that differ by ~1e-5 can legitimately reorder
    near-ties, so proposals are matched by box and the matched fraction must be (almost) everything."""
    H, W = 128, 160
    x = orc.make_image(H, W, seed=2)
    info = np.array([[H, W]], np.int32)
    eng = _engine(params, "bf16x3")
    prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
    cls_ref, box_ref, aux = orc.faster_rcnn_forward(x, params, info)
    rois_dev = plan.prop.rois.cpu().numpy()[: prob.shape[0]]
    rois_ref = aux["proposals"]
    # match proposals (same box within 1e-4 of the image scale)
    d = np.abs(rois_dev[:, None, :] - rois_ref[None, :, :]).max(-1)
    j = d.argmin(1)
    ok = d[np.arange(len(j)), j] < 1e-4 * max(H, W)
    assert ok.mean() > 0.97, ok.mean()
    p, b = prob.cpu().numpy()[ok], boxes.cpu().numpy()[ok]
    _check_probs(p, cls_ref[j[ok]])
    # pure end to end the decode multiplies the (<= 3e-5 relative) delta error by the box size: 3e-4 of the image
    # scale here (observed 2.2e-4); with identical inputs to the tail the boxes are bit-exact (stage-wise test)
    assert np.abs(b - box_ref[j[ok]]).max() < 3e-4 * max(H, W)


def test_graph_replay_is_deterministic_and_matches_eager(params):
    H, W = 96, 128
    eng_g = _engine(params, "bf16x3", use_graph=True)
    eng_e = _engine(params, "bf16x3", use_graph=False)
    for seed in (3, 4, 3):
        x = torch.from_numpy(orc.make_image(H, W, seed=seed)[0]).cuda()
        pg, bg, _ = eng_g(x)
        pe, be, _ = eng_e(x)
        assert pg.shape == pe.shape and torch.equal(pg, pe) and torch.equal(bg, be)


def test_stream_runner_host_to_host_matches_direct_forward(params):
    """The pipelined host->host API must return, for every image, exactly what a direct forward returns."""
    from frcnn_b200.engine import StreamRunner
    H, W = 96, 128
    eng = _engine(params, "bf16x3")
    plan = eng.plan(H, W, keep_rpn_debug=True)
    imgs = [torch.from_numpy(orc.make_image(H, W, seed=20 + i)[0]).pin_memory() for i in range(5)]
    want = []
    for im in imgs:
        p, b, c = plan.forward(im.cuda())
        torch.cuda.synchronize()
        want.append((p.cpu().clone(), b.cpu().clone(), int(c.item())))
    got = []
    runner = StreamRunner(plan)
    counts = runner.run(imgs, on_result=lambda i, r: got.append((i, r["prob"].copy(), r["boxes"].copy(), r["count"], r["rois"].copy())))
    assert counts == [w[2] for w in want] and [g[0] for g in got] == list(range(5))
    for (i, p, b, c, rois), (wp, wb, wc) in zip(got, want):
        # one D2H of the result block per image: rows [0, count) of (prob, boxes, proposals)
        assert c == wc and p.shape == (wc, 21) and np.array_equal(p, wp.numpy()[:wc]) and np.array_equal(b, wb.numpy()[:wc])
        assert rois.shape == (wc, 4)
    # numpy sources (pageable) and library-pinned blocks are accepted as well
    from frcnn_b200 import ops as _ops
    blk = _ops.PinnedBlock((3, H, W), np.float32)
    blk.np[...] = imgs[2].numpy()
    got2 = []
    runner.run([imgs[2].numpy(), blk], on_result=lambda i, r: got2.append(r["prob"].copy()))
    assert np.array_equal(got2[0], want[2][0].numpy()[:want[2][2]]) and np.array_equal(got2[1], got2[0])


def test_lanes_in_flight_are_bit_identical_to_single_lane(params):
    """3 images in flight on 3 streams (LanePool) and the lane-aware StreamRunner return, per image, exactly the
    single-lane result (no cross-lane state: the library keeps none, each lane has its own buffers and graph)."""
    from frcnn_b200.engine import LanePool, StreamRunner
    H, W = 96, 128
    eng = _engine(params, "bf16x3")
    plan = eng.plan(H, W)
    imgs = [torch.from_numpy(orc.make_image(H, W, seed=40 + i)[0]).pin_memory() for i in range(7)]
    dev = [im.cuda() for im in imgs]
    want = []
    for im in dev:
        p, b, c = plan.forward(im)
        torch.cuda.synchronize()
        want.append((p.cpu().clone(), b.cpu().clone(), int(c.item())))
    pool = LanePool(plan, lanes=3)
    assert len(pool) == 3 and pool.plans[0] is plan
    for rnd in range(0, 7, 3):                       # one round = one image per lane, then read the lanes back
        pool.fork()
        used = [(i, pool.submit(i, dev[i])) for i in range(rnd, min(rnd + 3, 7))]
        pool.join()
        torch.cuda.synchronize()
        for i, pl in used:
            assert int(pl.prop.count.item()) == want[i][2]
            assert torch.equal(pl.prob.cpu(), want[i][0]) and torch.equal(pl.boxes.cpu(), want[i][1])
    got = []
    runner = StreamRunner(pool, depth=4)
    counts = runner.run(imgs, on_result=lambda i, r: got.append((i, r["prob"].copy(), r["boxes"].copy())))
    assert counts == [w[2] for w in want] and [g[0] for g in got] == list(range(7))
    for (i, p, b), (wp, wb, wc) in zip(got, want):
        assert np.array_equal(p, wp.numpy()[:wc]) and np.array_equal(b, wb.numpy()[:wc])


def test_bf16_fast_mode_runs_and_is_close(params):
    """Single-pass bf16: same graph, lo planes absent.  Not the parity mode -- only sanity-checked."""
    H, W = 96, 128
    x = orc.make_image(H, W, seed=5)
    eng = _engine(params, "bf16")
    prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
    feat_ref = orc.vgg16_forward(x, params)
    assert _rel(plan.acts[-1].to_chw_f32().cpu().numpy()[None], feat_ref) < 3e-2
    assert prob.shape[1] == 21 and prob.shape[0] > 0
    np.testing.assert_allclose(prob.sum(1).cpu().numpy(), 1.0, rtol=1e-5)


def test_headline_config_600x1000(params):
    """BASELINE config #2 at full size: 600x1000, 300 proposals.  Size-independent properties, exact ProposalLayer
    parity on the device's own RPN outputs, detect vs the oracle, and the float tolerances at this size."""
    H, W = 600, 1000
    x = orc.make_image(H, W, seed=0)
    info = np.array([[H, W]], np.int32)
    eng = _engine(params, "bf16x3", with_detect=True, det_conf=0.05)
    prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
    R = prob.shape[0]
    assert (plan.fh, plan.fw) == (38, 63) and 0 < R <= 300
    rpn = plan.rpn_out.cpu().numpy()
    logits = rpn[:, :18].T.reshape(1, 18, 38, 63)
    deltas = rpn[:, 18:54].T.reshape(1, 36, 38, 63)
    probs = orc.softmax_axis1(logits)
    fg = probs[0, 9:].ravel()
    print("unique fg scores: %d of %d" % (np.unique(fg).size, fg.size))
    want_rois, want_fg = orc.proposal_layer(probs, deltas, info)
    assert R == len(want_rois)
    assert np.array_equal(plan.prop.rois.cpu().numpy()[:R], want_rois)
    assert np.array_equal(plan.prop.scores.cpu().numpy()[:R], want_fg.ravel())
    p, b = prob.cpu().numpy(), boxes.cpu().numpy()
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-5)
    assert b.min() >= 0 and b[:, 0::4].max() <= W - 1 and b[:, 1::4].max() <= H - 1
    keep_idx, keep_count, conf_count = [t.cpu().numpy() for t in plan.det]
    for c, keep, dets in orc.detect(p, b, 0.3, 0.05):
        assert keep_idx[c - 1, :conf_count[c - 1]].tolist() == keep.tolist()
    # float parity AT the headline size (the north star's 1e-4): trunk vs the fp32 oracle, head on the device's own
    # feature map and RoIs (RoI pool exact), class probabilities and boxes
    feat_dev = plan.acts[-1].to_chw_f32().cpu().numpy()[None]
    e_feat = _rel(feat_dev, orc.vgg16_forward(x, params))
    rois_dev = plan.prop.rois.cpu().numpy()[:R]
    cls_ref, box_ref, aux = orc.head_forward(feat_dev, rois_dev, params, info)
    pool_dev = (plan.pool5.hi.float() + plan.pool5.lo.float()).cpu().numpy().reshape(-1, 7, 7, 512)[:R]
    assert np.array_equal(pool_dev.transpose(0, 3, 1, 2), aux["pool5"])
    e_fc7 = _rel((plan.fc7.hi.float() + plan.fc7.lo.float()).cpu().numpy()[0, :R], aux["fc7"])
    e_box = _rel(b, box_ref)
    print("600x1000: conv5_3 %.2e, fc7 %.2e, boxes %.2e of max-norm vs the fp32 oracle" % (e_feat, e_fc7, e_box))
    assert e_feat < 1e-4 and e_fc7 < 1e-4 and e_box < 1e-4
    _check_probs(p, cls_ref)


# ------------------------------------------------------------------ edge cases of the whole graph and error paths
def test_engine_zero_proposals_and_tiny_image(params):
    """min_size larger than the image: the ProposalLayer keeps nothing; every later stage must cope with R == 0 (empty
    outputs, zero buffers, a replayable graph).  And the smallest legal image: one feature-map cell."""
    H, W = 96, 128
    x = orc.make_image(H, W, seed=9)
    for use_graph in (False, True):
        eng = _engine(params, "bf16x3", use_graph=use_graph, with_detect=True, min_size=10000)
        prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
        prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())               # second call: graph replay
        assert prob.shape == (0, 21) and boxes.shape == (0, 84) and int(plan.prop.count.item()) == 0
        assert not plan.prob.any().item() and not plan.boxes.any().item()
        assert not plan.det[1].any().item()                                  # keep_count of every class is 0
    want_rois, _ = orc.proposal_layer(np.full((1, 18, 6, 8), 1 / 18, f32), np.zeros((1, 36, 6, 8), f32), (H, W), min_size=10000)
    assert len(want_rois) == 0
    xs = orc.make_image(16, 16, seed=1)                                       # 16x16 -> a 1x1 feature map
    eng = _engine(params, "bf16x3", use_graph=False)
    prob, boxes, plan = eng(torch.from_numpy(xs[0]).cuda())
    assert (plan.fh, plan.fw) == (1, 1)
    feat_ref = orc.vgg16_forward(xs, params)
    assert _rel(plan.acts[-1].to_chw_f32().cpu().numpy()[None], feat_ref) < 1e-4
    rpn = plan.rpn_out.cpu().numpy()
    want_rois, _ = orc.proposal_layer(orc.softmax_axis1(rpn[:, :18].T.reshape(1, 18, 1, 1)), rpn[:, 18:54].T.reshape(1, 36, 1, 1),
                                      np.array([[16, 16]], np.int32))
    R = prob.shape[0]
    assert R == len(want_rois) and np.array_equal(plan.prop.rois.cpu().numpy()[:R], want_rois)


def test_portrait_maximum_size_1000x600(params):
    """The other orientation of the reference's size rule (forward.py:34-45: longest side capped at 1000): 1000 x 600."""
    H, W = 1000, 600
    x = orc.make_image(H, W, seed=3)
    eng = _engine(params, "bf16x3", use_graph=True)
    prob, boxes, plan = eng(torch.from_numpy(x[0]).cuda())
    assert (plan.fh, plan.fw) == (63, 38) and 0 < prob.shape[0] <= 300
    rpn = plan.rpn_out.cpu().numpy()
    logits = rpn[:, :18].T.reshape(1, 18, 63, 38)
    deltas = rpn[:, 18:54].T.reshape(1, 36, 63, 38)
    want_rois, want_fg = orc.proposal_layer(orc.softmax_axis1(logits), deltas, np.array([[H, W]], np.int32))
    R = prob.shape[0]
    assert R == len(want_rois) and np.array_equal(plan.prop.rois.cpu().numpy()[:R], want_rois)      # bit-exact at full size
    np.testing.assert_allclose(prob.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    b = boxes.cpu().numpy()
    assert (b[:, 0::4] >= 0).all() and (b[:, 2::4] <= W - 1).all() and (b[:, 1::4] >= 0).all() and (b[:, 3::4] <= H - 1).all()


def test_argument_errors_raise_instead_of_crashing(params):
    """The C ABI reports bad arguments through its status code + frcnn_last_error(); the Python layer raises FrcnnError."""
    from frcnn_b200 import ops
    from frcnn_b200._lib import FrcnnError
    x = ops.Act(torch.zeros((8, 8, 64), dtype=torch.bfloat16, device="cuda"), torch.zeros((8, 8, 64), dtype=torch.bfloat16, device="cuda"))
    w = torch.zeros((64, 64, 3, 3), device="cuda")
    hi, lo = ops.pack_conv_weights(w, cin_pad=64)
    b = ops.pad_bias(torch.zeros(64, device="cuda"), 64)
    with pytest.raises(FrcnnError):
        ops.conv2d(x, hi, lo, b, 1, True)                                     # 9 taps of weights for a 1x1 call
    with pytest.raises(FrcnnError):
        ops.conv2d(x, hi, None, b, 3, True)                                   # precision modes differ
    with pytest.raises(FrcnnError):
        ops.conv2d(x, hi, lo, torch.zeros(8, device="cuda"), 3, True)         # bias too short
    x48 = ops.Act(torch.zeros((8, 8, 48), dtype=torch.bfloat16, device="cuda"), torch.zeros((8, 8, 48), dtype=torch.bfloat16, device="cuda"))
    h48, l48 = ops.pack_conv_weights(torch.zeros((64, 48, 3, 3), device="cuda"), cin_pad=48)
    with pytest.raises(FrcnnError, match="Cin"):
        ops.conv2d(x48, h48, l48, b, 3, True)                                 # 32 < Cin < 64 is not a supported K tiling
    with pytest.raises(FrcnnError):
        ops.pack_image(torch.zeros((3, 8, 8)), c_pad=16)                      # host tensor: the library has no CPU path
    with pytest.raises(FrcnnError):
        ops.nms(torch.zeros((4, 5)), 0.7)
    # the library is still usable afterwards
    y, _ = ops.conv2d(x, hi, lo, b, 3, True)
    assert y.hi.shape == (8, 8, 64)


@pytest.mark.parametrize("precision,shape", [("bf16x3", (150, 201)), ("bf16", (96, 128)), ("bf16x3", (600, 1000))])
def test_single_c_abi_call_equals_the_engine(params, precision, shape):
    """frcnn_forward_vgg16 (one C call, caller-owned workspace) returns bit for bit what the Python-composed graph returns --
    it enqueues the same kernels in the same order -- including the Q7 clip bounds (img_info = (H, H))."""
    from frcnn_b200.engine import CForward
    H, W = shape
    eng = _engine(params, precision, use_graph=False)
    x = torch.from_numpy(orc.make_image(H, W, seed=11)[0]).cuda()
    for info in ((H, W), (H, H)):
        prob, boxes, plan = eng(x, img_info=info)
        cf = CForward(eng.weights, H, W, ANCHORS)
        p2, b2, c2 = cf.forward(x, im_info=info)
        torch.cuda.synchronize()
        R = int(c2.item())
        assert R == prob.shape[0] and torch.equal(p2[:R], prob) and torch.equal(b2[:R], boxes)
        assert not p2[R:].any().item() and not b2[R:].any().item()
    # the workspace query rejects a bad configuration through the error channel
    from frcnn_b200 import _lib
    import ctypes
    bad = _lib.ForwardConfig(8, 8, 21, 9, 16, 6000, 300, 16, 0.7, 1)
    assert _lib.load().frcnn_forward_workspace_bytes(ctypes.byref(bad)) == 0 and "too small" in _lib.last_error()
