class MaxNLocator:
    def __init__(self, *a, **k):
        pass
