class AttributeDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v
