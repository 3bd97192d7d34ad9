class ModelSummary:
    def __init__(self, *a, **k):
        pass
