class TestFailure(Exception):
    pass
