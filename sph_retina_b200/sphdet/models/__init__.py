"""Mirror of the reference's ``sphdet.models`` for the pieces of the heads that sit on the IoU hot path."""
