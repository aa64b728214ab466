"""``uresnet``: host-side mirror of /root/reference/lib/uresnet.py:13-123.

Same constructor (``dims, num_class, num_strides=5, base_num_outputs=16, debug=False``).
``_build`` does not create a TensorFlow graph; it returns the geometry from which the
C library derives the identical layer schedule (csrc/net.cu restates uresnet.py:22-123 and
resnet_module.py:10-87) and, with ``debug=True``, prints the same per-stage shapes the
reference prints (uresnet.py:24,53,63,90,94,101,112,122).
"""
from __future__ import annotations

from .ssnet import ssnet_base


class uresnet(ssnet_base):

    def __init__(self, dims, num_class, num_strides=5, base_num_outputs=16, debug=False):
        super(uresnet, self).__init__(dims=dims, num_class=num_class)
        self._base_num_outputs = int(base_num_outputs)
        self._num_strides = int(num_strides)
        self._debug = bool(debug)

    def _build(self, input_tensor=None):
        dims = [int(d) for d in self._dims]
        spatial, cin = dims[:-1], dims[-1]
        nd = len(spatial)
        F, S = self._base_num_outputs, self._num_strides
        if S > 5:
            # decoder scopes are hard-coded resnet_module(step+5) (uresnet.py:100): TF raises a
            # duplicate-variable error for num_strides > 5
            raise ValueError("Variable UResNet/resnet_module5/module1/shortcut/weights already exists "
                             "(num_strides=%d > 5: decoder scopes are step+5, uresnet.py:100)" % S)
        for d in spatial:
            if d % (1 << S):
                raise ValueError("spatial size %d is not divisible by 2^%d: the concat at "
                                 "uresnet.py:91 needs matching shapes" % (d, S))
        if self._debug:
            def shp(sp, c):
                return "(?, " + ", ".join(str(v) for v in sp) + ", %d)" % c
            print(shp(spatial, cin), 'input shape')
            print(shp(spatial, F), 'after conv0')
            sp, c = list(spatial), F
            for step in range(S):
                sp, c = [v // 2 for v in sp], c * 2
                print(shp(sp, c), 'after resnet_module%d' % step)
            for step in range(S):
                sp, c = [v * 2 for v in sp], c // 2
                print(shp(sp, c), 'after deconv%d' % step)
                print(shp(sp, 2 * c), 'after concat%d' % step)
                print(shp(sp, c), 'after resnet_module%d' % (step + S))
            print(shp(spatial, F), 'after conv1')
            print(shp(spatial, self._num_class), 'after conv2')
        return dict(nd=nd, spatial=spatial, in_channels=cin, base_filters=F, num_strides=S)
