"""models.vgg16.{VGG16, VGG16Prev} -- the two trunk classes of /root/reference models/vgg16.py:12-82.

Both are the 13-convolution VGG16 trunk up to the post-ReLU conv5_3 (no pool5): 3x3 stride-1 pad-1
convolutions with bias, ReLU after each, 2x2/2 ceil-mode max-pools after conv1_2, conv2_2, conv3_3
and conv4_3.  Link names are the reference's (conv1_1 ... conv5_3) so checkpoints load by path.
`VGG16` in the reference derives from chainer's VGG16Layers and would download a caffemodel; here
both classes start from random weights and differ only in name.

__call__(x) runs on the GPU: frcnn_pack_image_im2col3x3, 13x frcnn_conv2d (tcgen05) with the 4 ceil-mode pools fused into the epilogue.
"""
import numpy as np
import torch

from frcnn_b200 import arrays, links, ops

_PLAN = [("conv1_1", 3, 64), ("conv1_2", 64, 64), None,
         ("conv2_1", 64, 128), ("conv2_2", 128, 128), None,
         ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), None,
         ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), None,
         ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]


class VGG16Prev(links.Link):
    precision = "bf16x3"

    def __init__(self, train=False):
        super(VGG16Prev, self).__init__()
        self.__dict__["train"] = train
        for item in _PLAN:
            if item is not None:
                name, cin, cout = item
                # Chainer v1's default Convolution2D initialiser is un-vendored; LeCun-normal
                # (std = sqrt(1/fan_in)) is what v1.22 shipped.
                self.add_link(name, links.conv_link(cin, cout, 3, np.sqrt(1.0 / (9 * cin))))
        self.__dict__["_packed"] = (None, -1)

    def _weights(self, device):
        packed, ver = self._packed
        if packed is None or ver != self.version_key():
            packed = {}
            for item in _PLAN:
                if item is None:
                    continue
                name, cin, _ = item
                l = getattr(self, name)
                w = torch.from_numpy(l.W.data).to(device)
                if cin <= 3:      # first layer: K = 27 (+5 zeros) GEMM over the im2col-packed image
                    hi, lo = ops.pack_conv_weights_im2col(w, precision=self.precision)
                else:
                    hi, lo = ops.pack_conv_weights(w, cin_pad=cin, precision=self.precision)
                packed[name] = (hi, lo, ops.pad_bias(torch.from_numpy(l.b.data).to(device), l.b.data.size))
            self.__dict__["_packed"] = (packed, self.version_key())
        return packed

    def forward_device(self, x_chw):
        """(3,H,W) CUDA float32 -> ops.Act [h,w,512] (NHWC bf16 hi/lo)."""
        packed = self._weights(x_chw.device)
        act = ops.pack_image_im2col(x_chw, precision=self.precision)
        for i, item in enumerate(_PLAN):
            if item is None:
                continue                      # the 2x2 ceil-mode pool is fused into the preceding conv's epilogue
            hi, lo, b = packed[item[0]]
            pooled = i + 1 < len(_PLAN) and _PLAN[i + 1] is None
            act, _ = ops.conv2d(act, hi, lo, b, 1 if i == 0 else 3, True, fuse_pool=pooled)
        return act

    def __call__(self, x):
        from chainer import Variable
        fam = arrays.family(x)
        t = arrays.to_device(x)
        if t.dim() != 4 or t.shape[0] != 1:
            raise ValueError("trunk expects a (1, 3, H, W) batch (the reference asserts batch size 1)")
        feat = self.forward_device(t[0]).to_chw_f32()[None]
        return Variable(arrays.from_device(feat, fam))


class VGG16(VGG16Prev):
    def __init__(self):
        super(VGG16, self).__init__(train=True)
