"""``fireflies/entity/curve.py`` -- NURBS-curve camera paths.  Dead in the reference (the constructor raises,
entity/curve.py:24; the arithmetic lives in the un-vendored geomdl 5.3.1) and outside the hot path
(SURVEY.md section 2 #3b): kept as an explicit stub so imports do not break."""


class Curve:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "fireflies_b200: Curve (geomdl NURBS camera paths) is outside the B200 hot path; "
            "the reference's own Curve constructor raises as well (entity/curve.py:24)")
