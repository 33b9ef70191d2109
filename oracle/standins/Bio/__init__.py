"""Stand-in for the parts of biopython==1.84 the reference touches."""


class BiopythonWarning(Warning):
    pass
