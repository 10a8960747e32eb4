class Ellipse:
    pass


class Polygon:
    pass
