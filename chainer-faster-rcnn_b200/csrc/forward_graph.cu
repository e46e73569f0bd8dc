// forward_graph.cu -- frcnn_forward_vgg16: the whole forward detection path behind ONE C-ABI call.
//
// FasterRCNN.__call__, inference branch (/root/reference models/faster_rcnn.py:92-134,175-178): trunk (models/vgg16.py:38-82)
// -> RPN (models/region_proposal_network.py:117-124) -> ProposalLayer (models/proposal_layer.py:102-198) -> RoI pooling ->
// fc6 / fc7 -> cls_score | bbox_pred -> softmax + decode + clip.  This is pure host orchestration: it carves the caller's
// workspace and enqueues the same kernels, in the same order, as frcnn_b200/engine.py -- for callers that have no Python
// (the launch sequence is static, so the caller may capture it into a CUDA graph).  No allocation, no synchronisation.
#include "common.cuh"

using namespace frcnn;

namespace {

struct Plane { void *hi, *lo; };

struct Carver {
    char* base;
    size_t off;
    bool x3;
    void* take(size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    }
    Plane act(long elems) {
        Plane p;
        p.hi = take((size_t)elems * 2);
        p.lo = x3 ? take((size_t)elems * 2) : nullptr;
        return p;
    }
};

// VGG16 trunk table (models/vgg16.py:38-69): {Cin, Cout, pool after}
const int kVgg[13][3] = {{3, 64, 0},    {64, 64, 1},   {64, 128, 0},  {128, 128, 1}, {128, 256, 0}, {256, 256, 0}, {256, 256, 1},
                         {256, 512, 0}, {512, 512, 0}, {512, 512, 1}, {512, 512, 0}, {512, 512, 0}, {512, 512, 0}};

struct Layout {
    Plane x_col, act[13], rpn_mid, pool5, fc6, fc7;
    float *rpn_out, *head_out, *rois, *scores;
    void *prop_ws, *fc_ws;
    size_t prop_ws_bytes, fc_ws_bytes;
    int fh, fw, rpn_ld, head_ld;
    size_t total;
};

int round32(int v) { return (v + 31) / 32 * 32; }

Layout carve(const frcnn_forward_config& c, void* workspace) {
    Layout L;
    Carver k{(char*)workspace, 0, c.x3 != 0};
    L.x_col = k.act((long)frcnn_image_c8_elems(c.H, c.W));          // compact first-layer input [H][W+2][8]
    int h = c.H, w = c.W;
    for (int i = 0; i < 13; ++i) {
        if (kVgg[i][2]) { h = (h + 1) / 2; w = (w + 1) / 2; }          // the pool is fused into this conv's epilogue
        L.act[i] = k.act((long)h * w * kVgg[i][1]);
    }
    L.fh = h; L.fw = w;
    L.rpn_ld = round32(6 * c.n_anchors);
    L.head_ld = round32(5 * c.num_classes);
    L.rpn_mid = k.act((long)h * w * 512);
    L.rpn_out = (float*)k.take((size_t)h * w * L.rpn_ld * 4);
    L.prop_ws_bytes = frcnn_proposals_workspace_bytes(c.n_anchors, h, w, c.pre_nms_top_n);
    L.prop_ws = k.take(L.prop_ws_bytes);
    L.rois = (float*)k.take((size_t)c.post_nms_top_n * 4 * 4);
    L.scores = (float*)k.take((size_t)c.post_nms_top_n * 4);
    L.pool5 = k.act((long)c.post_nms_top_n * 49 * 512);
    L.fc6 = k.act((long)c.post_nms_top_n * 4096);
    L.fc7 = k.act((long)c.post_nms_top_n * 4096);
    L.head_out = (float*)k.take((size_t)c.post_nms_top_n * L.head_ld * 4);
    L.fc_ws_bytes = frcnn_linear_workspace_bytes(c.post_nms_top_n, 49 * 512, 4096);
    const size_t b7 = frcnn_linear_workspace_bytes(c.post_nms_top_n, 4096, 4096);
    const size_t bh = frcnn_linear_workspace_bytes(c.post_nms_top_n, 4096, 5 * c.num_classes);
    if (b7 > L.fc_ws_bytes) L.fc_ws_bytes = b7;
    if (bh > L.fc_ws_bytes) L.fc_ws_bytes = bh;
    L.fc_ws = k.take(L.fc_ws_bytes);
    L.total = k.off;
    return L;
}

int check_config(const frcnn_forward_config* c) {
    FRCNN_REQUIRE(c != nullptr, "frcnn_forward: null config");
    FRCNN_REQUIRE(c->H >= 16 && c->W >= 16, "frcnn_forward: image %dx%d too small (one feature cell needs 16x16)", c->H, c->W);
    FRCNN_REQUIRE(c->num_classes >= 2 && c->n_anchors >= 1 && c->feat_stride == 16, "frcnn_forward: VGG16 trunk has feat_stride 16");
    FRCNN_REQUIRE(c->pre_nms_top_n > 0 && c->post_nms_top_n > 0 && c->post_nms_top_n <= 2048, "frcnn_forward: bad top-N limits");
    return FRCNN_OK;
}

}  // namespace

extern "C" {

size_t frcnn_forward_workspace_bytes(const frcnn_forward_config* config) {
    FRCNN_ENTRY();
    if (check_config(config) != FRCNN_OK) return 0;
    return carve(*config, nullptr).total;
}

int frcnn_forward_vgg16(const frcnn_forward_config* config, const frcnn_vgg16_weights* wts, const float* image_chw, int im_h,
                        int im_w, void* workspace, size_t workspace_bytes, float* out_prob, float* out_boxes, int* out_count,
                        void* stream) {
    FRCNN_ENTRY();
    int rc = check_config(config);
    if (rc != FRCNN_OK) return rc;
    const frcnn_forward_config& c = *config;
    FRCNN_REQUIRE(wts && image_chw && workspace && out_prob && out_boxes && out_count, "frcnn_forward: null pointer");
    const Layout L = carve(c, workspace);
    if (workspace_bytes < L.total) {
        set_error("frcnn_forward: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return FRCNN_ERR_WORKSPACE;
    }
    const bool x3 = c.x3 != 0;
    for (int i = 0; i < 13; ++i)
        FRCNN_REQUIRE(wts->conv[i].hi && wts->conv[i].bias && (!x3 || wts->conv[i].lo), "frcnn_forward: trunk layer %d weights missing", i);
    // ---- trunk: conv1_1 as a K = 3 x 32 GEMM over the compact image (sliding-window tensor map), then 12 shared-halo 3x3 convs
    if ((rc = frcnn_pack_image_c8(image_chw, 3, c.H, c.W, (long)c.H * c.W, c.W, 1, L.x_col.hi, L.x_col.lo, stream)) != FRCNN_OK) return rc;
    Plane x = L.x_col;
    int h = c.H, w = c.W, cin = 32;
    for (int i = 0; i < 13; ++i) {
        const frcnn_packed_layer& l = wts->conv[i];
        if (i == 0)
            rc = frcnn_conv3x3_c8(x.hi, x.lo, h, w, l.hi, l.lo, l.bias, kVgg[0][1], 1, L.act[0].hi, L.act[0].lo, stream);
        else
            rc = frcnn_conv2d(x.hi, x.lo, h, w, cin, l.hi, l.lo, l.bias, kVgg[i][1], 3, 1, kVgg[i][2], L.act[i].hi, L.act[i].lo, nullptr,
                              0, nullptr, stream);
        if (rc != FRCNN_OK) return rc;
        if (kVgg[i][2]) { h = (h + 1) / 2; w = (w + 1) / 2; }
        x = L.act[i];
        cin = kVgg[i][1];
    }
    const Plane feat = x;
    // ---- RPN: 3x3 conv + ReLU, the merged twin 1x1 heads (fp32 [H*W, rpn_ld]), ProposalLayer with the 2A-way softmax fused
    if ((rc = frcnn_conv2d(feat.hi, feat.lo, h, w, 512, wts->rpn3.hi, wts->rpn3.lo, wts->rpn3.bias, 512, 3, 1, 0, L.rpn_mid.hi,
                           L.rpn_mid.lo, nullptr, 0, nullptr, stream)) != FRCNN_OK) return rc;
    if ((rc = frcnn_conv2d(L.rpn_mid.hi, L.rpn_mid.lo, h, w, 512, wts->rpn_heads.hi, wts->rpn_heads.lo, wts->rpn_heads.bias,
                           6 * c.n_anchors, 1, 0, 0, nullptr, nullptr, L.rpn_out, L.rpn_ld, nullptr, stream)) != FRCNN_OK) return rc;
    if ((rc = frcnn_proposals(L.rpn_out, 1, L.rpn_ld, 1, L.rpn_out + 2 * c.n_anchors, 1, L.rpn_ld, wts->anchors, c.n_anchors, h, w,
                              c.feat_stride, im_h, im_w, c.min_size, c.pre_nms_top_n, c.post_nms_top_n, c.nms_thresh, L.rois,
                              L.scores, out_count, nullptr, nullptr, nullptr, L.prop_ws, L.prop_ws_bytes, stream)) != FRCNN_OK) return rc;
    // ---- RoI pooling + head (rows past *out_count are zero)
    const int R = c.post_nms_top_n;
    if ((rc = frcnn_roi_pool(feat.hi, feat.lo, h, w, 512, L.rois, out_count, R, 7, 7, 1.0f / (float)c.feat_stride, L.pool5.hi,
                             L.pool5.lo, nullptr, stream)) != FRCNN_OK) return rc;
    if ((rc = frcnn_linear(L.pool5.hi, L.pool5.lo, R, 49 * 512, wts->fc6.hi, wts->fc6.lo, wts->fc6.bias, 4096, 1, out_count, L.fc6.hi,
                           L.fc6.lo, nullptr, 0, L.fc_ws, L.fc_ws_bytes, stream)) != FRCNN_OK) return rc;
    if ((rc = frcnn_linear(L.fc6.hi, L.fc6.lo, R, 4096, wts->fc7.hi, wts->fc7.lo, wts->fc7.bias, 4096, 1, out_count, L.fc7.hi, L.fc7.lo,
                           nullptr, 0, L.fc_ws, L.fc_ws_bytes, stream)) != FRCNN_OK) return rc;
    if ((rc = frcnn_linear(L.fc7.hi, L.fc7.lo, R, 4096, wts->head.hi, wts->head.lo, wts->head.bias, 5 * c.num_classes, 0, out_count,
                           nullptr, nullptr, L.head_out, L.head_ld, L.fc_ws, L.fc_ws_bytes, stream)) != FRCNN_OK) return rc;
    return frcnn_head_decode(L.head_out, L.head_out + c.num_classes, L.head_ld, L.rois, out_count, R, c.num_classes, im_h, im_w,
                             out_prob, out_boxes, stream);
}

}  // extern "C"
