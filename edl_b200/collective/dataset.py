"""User-subclassed record iterators (reference: python/edl/collective/dataset.py:16-45)."""


class FileSplitter:
    """Yield the records of one file as tuples ``(record_no, field0, field1, ...)``."""

    def split(self, path):
        raise NotImplementedError

    def __call__(self, path):
        return self.split(path)


class TxtFileSplitter(FileSplitter):
    """One record per non-empty line: ``(line_no, line_text)``."""

    def split(self, path):
        with open(path, "r") as f:
            for i, line in enumerate(f):
                line = line.rstrip("\n")
                if line:
                    yield (i, line)
