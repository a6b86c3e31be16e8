"""``WeightedCrossentropy`` with the reference's name and config keys
(``annotator/utils/losses.py:40-84``; looked up by name from
``deploy_options.loss``, ``engine.py:270-271``).

The arithmetic (positive rate, weight mask, BCE-from-logits, its gradient) runs
inside ``dnnca_label_stats`` + ``dnnca_head_bce_fwd_bwd`` fused with the 1x1
sigmoid head; this class only carries the configuration.
"""
from __future__ import annotations

from .. import native as N


class WeightedCrossentropy:
    def __init__(self, weight=None, weight_add=0.0, weight_mul=1.0, label_smoothing=False,
                 label_smoothing_filter_size=6, label_smoothing_sigma=3, **kargs):
        kargs.pop('name', None)
        kargs.pop('reduction', None)
        if kargs:
            raise TypeError(f'unexpected loss config keys {sorted(kargs)}')
        self.weight = weight
        self.weight_add = weight_add
        self.weight_mul = weight_mul
        self.label_smoothing = label_smoothing
        self.label_smoothing_filter_size = label_smoothing_filter_size
        self.label_smoothing_sigma = label_smoothing_sigma
        self.name = 'weighted_crossentropy'

    def get_config(self):
        return dict(weight=self.weight, weight_add=self.weight_add, weight_mul=self.weight_mul,
                    label_smoothing=self.label_smoothing,
                    label_smoothing_filter_size=self.label_smoothing_filter_size,
                    label_smoothing_sigma=self.label_smoothing_sigma)

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    def prepare_labels(self, y):
        """losses.py:60-67: the labels the loss sees -- gaussian-filtered on the device when ``label_smoothing`` is set
        (``dnnca_gaussian_filter2d``), else ``y`` itself.  ``y``: ``[B,H,W]`` float32 CUDA tensor."""
        if not self.label_smoothing:
            return y
        import torch
        if not (torch.is_tensor(y) and y.is_cuda and y.dtype == torch.float32 and y.dim() == 3):
            raise ValueError('label smoothing expects a [B,H,W] float32 CUDA tensor')
        y = y.contiguous()
        tmp, out = torch.empty_like(y), torch.empty_like(y)
        N.call('dnnca_gaussian_filter2d', N.stream_ptr(), N.ptr(y), y.shape[0], y.shape[1], y.shape[2],
               int(self.label_smoothing_filter_size), float(self.label_smoothing_sigma), N.ptr(tmp), N.ptr(out))
        return out

    def native_config(self, numel_times_replicas) -> N.LossConfig:
        """``dnnca_loss_config_t``; grad_scale = 1/(B*H*W*replicas): mean over H,W (losses.py:36),
        keras' mean over the batch and the 1/replicas of MirroredStrategy."""
        return N.LossConfig(float(self.weight) if self.weight is not None else 0.0,
                            1 if self.weight is not None else 0, float(self.weight_add), float(self.weight_mul),
                            1.0 / float(numel_times_replicas))


TFWeightedCrossentropy = WeightedCrossentropy   # the reference's class name (losses.py:40)

_REGISTRY = {'WeightedCrossentropy': WeightedCrossentropy, 'weighted_crossentropy': WeightedCrossentropy}


def get(identifier):
    """``tf.keras.losses.get`` for the names the reference registers (losses.py:105-106)."""
    if isinstance(identifier, WeightedCrossentropy):
        return identifier
    if isinstance(identifier, str):
        if identifier in _REGISTRY:
            return _REGISTRY[identifier]()
        raise ValueError(f'unknown loss {identifier!r}')
    if isinstance(identifier, dict):
        cls = identifier.get('class_name')
        if cls in _REGISTRY:
            return _REGISTRY[cls](**(identifier.get('config') or {}))
        raise ValueError(f'unknown loss {cls!r}')
    raise ValueError(f'cannot interpret loss identifier {identifier!r}')
