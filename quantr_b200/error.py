"""QuantrError: mirror of src/error.rs:17-33."""


class QuantrError(Exception):
    def __init__(self, message: str):
        super().__init__(message)
        self.message = message

    def __str__(self):  # error.rs:21-25 (ANSI red)
        return f"\x1b[91m[Quantr Error] {self.message}\x1b[0m "
