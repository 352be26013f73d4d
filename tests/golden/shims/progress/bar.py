class ShadyBar:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def next(self):
        pass

    def finish(self):
        pass
