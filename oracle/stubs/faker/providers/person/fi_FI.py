"""Import stub for the reference engine (main.pyx:13,163-164 only uses the names in debug strings)."""


class Provider:
    first_names = {'A': 1}
    last_names = {'B': 1}
