class BatchedData:
    """type bound only (models/mattergen/pl_module.py:13, loss.py:8)"""
