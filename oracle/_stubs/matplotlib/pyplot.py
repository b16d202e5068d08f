class _Style:
    def use(self, *a, **k):
        pass


style = _Style()
