"""Import stub for the reference engine (main.pyx:15 imports Intervention; only dead make_iv uses it).

Carries just what Context.apply_intervention reads: .type, .date, .get_param_values()
(common/interventions.py:59-120).
"""


class Intervention:
    def __init__(self, type, date=None, values=None):
        self.type = type
        self.date = date
        self.values = dict(values or {})

    def get_param_values(self):
        return dict(self.values)
