from .detection_collate_fn import DatasetItemsException, DetectionCollateFN  # noqa: F401
