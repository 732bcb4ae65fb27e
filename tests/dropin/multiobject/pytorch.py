class MultiObjectDataset:
    pass


class MultiObjectDataLoader:
    pass
