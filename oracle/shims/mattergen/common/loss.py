class MaterialsLoss:
    """constructor contract of mattergen.common.loss.MaterialsLoss as far as SampleLoss relies on it: `loss_fns` (one
    per-field callable for every included field, here placeholders the tests replace) and `loss_weights`"""

    def __init__(self, reduce="mean", d3pm_hybrid_lambda=0.0, include_pos=True, include_cell=True, include_atomic_numbers=True,
                 weights=None):
        self.reduce, self.d3pm_hybrid_lambda = reduce, d3pm_hybrid_lambda
        self.loss_fns = {}
        if include_pos:
            self.loss_fns["pos"] = None
        if include_cell:
            self.loss_fns["cell"] = None
        if include_atomic_numbers:
            self.loss_fns["atomic_numbers"] = None
        self.loss_weights = weights
