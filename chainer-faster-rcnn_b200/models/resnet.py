"""models.resnet.ResNet -- the trunk class of /root/reference models/resnet.py:11-45 (a subclass of the un-vendored
chainer.links.model.vision.resnet.ResNetLayers that drops fc6/prob and returns 'res5' in test mode when `train` is False).

The reference constructor downloads a caffemodel (models/resnet.py:21-36, dead links, no network here); this class builds
the same link tree -- conv1/bn1, res2..res5 with blocks a, b1, b2, ... holding conv1..conv3(/conv4) and bn1..bn3(/bn4), BN
entries gamma / beta / avg_mean / avg_var -- with random initial values, so a converted checkpoint in the reference's
.npz naming loads by path (chainer.serializers.load_npz).  Forward = frcnn_b200.resnet_engine (every BN folded into its
convolution, tcgen05 kernels); `FasterRCNN(trunk_class=functools.partial(ResNet, 101), rpn_in_ch=2048, feat_stride=32)`
is the composition SURVEY.md 8f defines (the reference's own `trunk_class()` call passes no n_layers, faster_rcnn.py:29).
"""
import numpy as np

from frcnn_b200 import arrays, links
from frcnn_b200.resnet_engine import block_list


def _bn_link(c, rng, gamma=1.0):
    l = links.Link()
    l.add_param("gamma", gamma * np.ones(c))
    l.add_param("beta", np.zeros(c))
    l.add_param("avg_mean", np.zeros(c))
    l.add_param("avg_var", np.ones(c))
    return l


def _conv_link(cin, cout, k, rng, bias=False):
    l = links.Link()
    l.add_param("W", rng.normal(0.0, np.sqrt(2.0 / (cin * k * k)), size=(cout, cin, k, k)))
    if bias:
        l.add_param("b", np.zeros(cout))
    return l


class ResNet(links.Link):
    feat_stride = 32
    out_channels = 2048
    precision = "bf16x3"

    def __init__(self, n_layers, rng=np.random):
        super(ResNet, self).__init__()
        if n_layers not in (50, 101, 152):
            raise ValueError("n_layers must be 50, 101 or 152 (models/resnet.py:27-32)")
        self.__dict__["n_layers"] = n_layers
        self.add_link("conv1", _conv_link(3, 64, 7, rng, bias=True))
        self.add_link("bn1", _bn_link(64, rng))
        stages = {}
        for stage, blk, cin, mid, cout, stride, proj in block_list(n_layers):
            if stage not in stages:
                stages[stage] = links.Link()
                self.add_link(stage, stages[stage])
            b = links.Link()
            b.add_link("conv1", _conv_link(cin, mid, 1, rng))
            b.add_link("bn1", _bn_link(mid, rng))
            b.add_link("conv2", _conv_link(mid, mid, 3, rng))
            b.add_link("bn2", _bn_link(mid, rng))
            b.add_link("conv3", _conv_link(mid, cout, 1, rng))
            b.add_link("bn3", _bn_link(cout, rng, gamma=0.3))
            if proj:
                b.add_link("conv4", _conv_link(cin, cout, 1, rng))
                b.add_link("bn4", _bn_link(cout, rng))
            stages[stage].add_link(blk, b)
        self.__dict__["train"] = True              # models/resnet.py:41
        self.__dict__["_engine"] = (None, -1)

    def __call__(self, x):
        """(1,3,H,W) image -> Variable holding res5 (1,2048,H/32,W/32) (models/resnet.py:43-45, test-mode BN)."""
        from chainer import Variable
        fam = arrays.family(x)
        t = arrays.to_device(x)
        if t.dim() != 4 or t.shape[0] != 1:
            raise ValueError("trunk expects a (1, 3, H, W) batch (the reference asserts batch size 1)")
        feat = self.forward_device(t[0]).to_chw_f32()[None]
        return Variable(arrays.from_device(feat, fam))

    def forward_device(self, x_chw):
        """(3,H,W) CUDA float32 -> ops.Act [h,w,2048] through a trunk-only plan (no RPN / head weights needed)."""
        from frcnn_b200 import resnet_engine as re_
        pk, ver = self._engine
        if pk is None or ver != self.version_key():
            params = {"trunk/" + k.lstrip("/"): p.data for k, p in self.namedparams()}
            pk = re_.TrunkOnly(params, self.n_layers, self.precision, x_chw.device)
            self.__dict__["_engine"] = (pk, self.version_key())
        return pk.run(x_chw)
