"""Stub of the un-vendored `mattergen` package (microsoft/mattergen@5bb2b397, env.yml:31): ONLY the names the reference's
in-tree adapter (models/mattergen/{pl_module,loss}.py) imports, so that those two files import unmodified and their
own arithmetic — the fine-tune time grid, the per-sample loss aggregation, the KL proxy — can be pinned.  The score
network, the corruption processes and the per-field loss functions are NOT restated here: tests inject stand-ins.
Test infrastructure, never shipped."""
