def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name}: plotting is not available in the test harness")
