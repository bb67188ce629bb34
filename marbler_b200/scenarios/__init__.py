from .scenarios import PredatorCapturePrey, Warehouse, MaterialTransport, ArcticTransport, simple  # noqa: F401
