"""Engine layer: YOLOLite facade, DetectionPredictor, Results."""
