"""Minimal stand-in for `prettytable.PrettyTable` (reference: Modules.py:5,315-324)."""


class PrettyTable:
    def __init__(self, field_names=None):
        self.field_names = list(field_names or [])
        self.rows = []

    def add_row(self, row):
        self.rows.append([str(x) for x in row])

    def __str__(self):
        cols = [self.field_names] + self.rows
        w = [max(len(str(r[i])) for r in cols) for i in range(len(self.field_names))]
        line = "+" + "+".join("-" * (x + 2) for x in w) + "+"
        fmt = lambda r: "|" + "|".join(f" {str(c):<{w[i]}} " for i, c in enumerate(r)) + "|"
        return "\n".join([line, fmt(self.field_names), line] + [fmt(r) for r in self.rows] + [line])
