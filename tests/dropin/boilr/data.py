class BaseDatasetManager:
    """train / test loaders plus data_shape, img_size, color_ch."""
