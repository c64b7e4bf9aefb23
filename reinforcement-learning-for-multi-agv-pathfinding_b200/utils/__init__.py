"""Mirror of the reference's src/utils package (constants + the device-backed "A*")."""
