"""Mirror of the reference's `evaluations` package for the centerOffsetRes10 path."""
