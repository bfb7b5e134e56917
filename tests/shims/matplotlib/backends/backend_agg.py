from ..pyplot import _Canvas


class FigureCanvasAgg(_Canvas):
    pass
