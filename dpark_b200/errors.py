"""Exceptions of the shuffle path, named as in the reference."""


class DparkUserFatalError(Exception):
    """Raised for rows that are not (k, v) pairs -- dpark/task.py:216-219,
    dpark/utils/__init__.py DparkUserFatalError."""
