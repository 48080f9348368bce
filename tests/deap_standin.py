"""A functional stand-in for the handful of DEAP names the reference GA driver uses (deap 1.3.1 is not
installed in this image and there is no network).  TEST INFRASTRUCTURE ONLY: it lets the test-suite run the
reference's own, unmodified ``ruleMaker.__eaMuPlusLambdaAdaptive`` (ruleMaker.py:953-1075) so that the
"library swapped, Python unchanged" claim is exercised at the level of a whole GA run.

Covered (ruleMaker.py:21-64, 506-531, 957-1014): creator.create, base.Fitness (values / wvalues / valid / del /
dominates, crowding_dist attribute), base.Toolbox (register / clone), tools.initIterate / initRepeat /
Statistics / Logbook / cxTwoPoint / mutFlipBit."""
import copy
import random
import sys
import types
from functools import partial


class Fitness:
    weights = None

    def __init__(self, values=()):
        self.wvalues = ()
        if values:
            self.values = values

    def getValues(self):
        return tuple(w / wt for w, wt in zip(self.wvalues, self.weights))

    def setValues(self, values):
        self.wvalues = tuple(v * w for v, w in zip(values, self.weights))

    def delValues(self):
        self.wvalues = ()

    values = property(getValues, setValues, delValues)

    @property
    def valid(self):
        return len(self.wvalues) != 0

    def dominates(self, other):
        not_equal = False
        for a, b in zip(self.wvalues, other.wvalues):
            if a > b:
                not_equal = True
            elif a < b:
                return False
        return not_equal

    def __hash__(self):
        return hash(self.wvalues)

    def __eq__(self, other):
        return self.wvalues == other.wvalues

    def __deepcopy__(self, memo):
        c = self.__class__()
        c.wvalues = self.wvalues
        for k, v in self.__dict__.items():
            if k != "wvalues":
                setattr(c, k, copy.deepcopy(v, memo))
        return c


class Toolbox:
    def __init__(self):
        self.register("clone", copy.deepcopy)

    def register(self, alias, function, *args, **kargs):
        pfunc = partial(function, *args, **kargs)
        pfunc.__name__ = alias
        setattr(self, alias, pfunc)


class _Creator(types.ModuleType):
    def create(self, name, base, **kargs):
        inst_attrs = {k: v for k, v in kargs.items() if isinstance(v, type)}
        cls_attrs = {k: v for k, v in kargs.items() if not isinstance(v, type)}

        def init(self, *a, **kw):
            for k, v in inst_attrs.items():
                setattr(self, k, v())
            if base.__init__ is not object.__init__:
                base.__init__(self, *a, **kw)
        setattr(self, name, type(name, (base,), dict(cls_attrs, __init__=init)))


def initIterate(container, generator):
    return container(generator())


def initRepeat(container, func, n):
    return container(func() for _ in range(n))


class Statistics:
    def __init__(self, key=lambda x: x):
        self.key, self.functions, self.fields = key, {}, []

    def register(self, name, function, *args, **kargs):
        self.functions[name] = partial(function, *args, **kargs)
        self.fields.append(name)

    def compile(self, data):
        values = tuple(self.key(e) for e in data)
        return {k: f(values) for k, f in self.functions.items()}


class Logbook(list):
    def __init__(self):
        super().__init__()
        self.header = None

    def record(self, **infos):
        self.append(infos)

    @property
    def stream(self):
        return str(self[-1]) if self else ""

    def select(self, *names):
        if len(names) == 1:
            return [e.get(names[0]) for e in self]
        return tuple([e.get(n) for e in self] for n in names)


def cxTwoPoint(ind1, ind2):
    size = min(len(ind1), len(ind2))
    a, b = random.randint(1, size), random.randint(1, size - 1)
    if b >= a:
        b += 1
    else:
        a, b = b, a
    ind1[a:b], ind2[a:b] = ind2[a:b], ind1[a:b]
    return ind1, ind2


def mutFlipBit(individual, indpb):
    for i in range(len(individual)):
        if random.random() < indpb:
            individual[i] = type(individual[i])(not individual[i])
    return individual,


def install():
    """Registers the stand-in as ``deap`` (only if the real package is absent)."""
    try:
        import deap  # noqa: F401
        if not isinstance(sys.modules["deap"], types.ModuleType) or hasattr(sys.modules["deap"], "__scb_standin__"):
            raise ImportError
        return
    except ImportError:
        pass
    deap = types.ModuleType("deap")
    deap.__scb_standin__ = True
    base = types.ModuleType("deap.base")
    base.Fitness, base.Toolbox = Fitness, Toolbox
    creator = _Creator("deap.creator")
    tools = types.ModuleType("deap.tools")
    for f in (initIterate, initRepeat, Statistics, Logbook, cxTwoPoint, mutFlipBit):
        setattr(tools, f.__name__, f)
    gp = types.ModuleType("deap.gp")
    algorithms = types.ModuleType("deap.algorithms")
    deap.base, deap.creator, deap.tools, deap.gp, deap.algorithms = base, creator, tools, gp, algorithms
    for name, mod in (("deap", deap), ("deap.base", base), ("deap.creator", creator), ("deap.tools", tools),
                      ("deap.gp", gp), ("deap.algorithms", algorithms)):
        sys.modules[name] = mod
