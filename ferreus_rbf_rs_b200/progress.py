"""Mirror of the reference module ``ferreus_rbf.progress`` (py_ferreus_rbf/src/python_bindings.rs:282-397)."""


class SolverIteration:
    def __init__(self, iter, residual, progress):
        self.iter, self.residual, self.progress = iter, residual, progress

    def __repr__(self):
        return f"SolverIteration(iter={self.iter}, residual={self.residual:.3e}, progress={self.progress:.3f})"


class DuplicatesRemoved:
    def __init__(self, num_duplicates):
        self.num_duplicates = num_duplicates


class Message:
    def __init__(self, message):
        self.message = message


class Progress:
    """Progress(callback=None): the callback receives one event object per solver event; exceptions raised by
    the callback are printed, not propagated (python_bindings.rs:359-397)."""

    def __init__(self, callback=None):
        self.callback = callback

    def _emit(self, event):
        if self.callback is None:
            return
        try:
            self.callback(event)
        except Exception as exc:  # noqa: BLE001
            print(f"progress callback raised: {exc!r}")
