def img_grid_pad_value(imgs, thresh=0.2):
    return 0.5
