class mesh:
    class Mesh:
        pass
