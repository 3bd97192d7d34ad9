class AdvancedProfiler:
    def __init__(self, *a, **k):
        pass
