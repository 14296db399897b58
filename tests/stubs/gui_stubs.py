"""Stand-ins for the packages the reference GUI imports but this image does not have (flask,
flask_socketio, PyQt5, matplotlib, serial).  TEST INFRASTRUCTURE: they let tests import the
reference's scripts/fft_analyzer_gui.py - patched by fpga_real_time_fft_analyzer_b200.gui_patch -
and drive its ReceiverController / UdpReceiver for real, with socket.io emits recorded instead
of sent and Qt's event loop replaced by direct calls."""
from __future__ import annotations

import sys
import time
import types

EMITTED = []            # (event, payload) of every socketio.emit
TIMERS = []             # live QTimer stubs; tests fire them by hand


def install():
    """Put the stub modules into sys.modules (idempotent)."""
    if "flask_socketio" in sys.modules and getattr(sys.modules["flask_socketio"], "_fra_stub", False):
        return
    # ---- flask
    flask = types.ModuleType("flask")

    class Flask:
        def __init__(self, *a, **k):
            self.config = {}

        def route(self, *a, **k):
            return lambda f: f
    flask.Flask = Flask
    flask.render_template = lambda *a, **k: ""
    sys.modules["flask"] = flask
    # ---- flask_socketio
    fsio = types.ModuleType("flask_socketio")
    fsio._fra_stub = True

    class SocketIO:
        def __init__(self, *a, **k):
            self.handlers = {}

        def on(self, event):
            def deco(f):
                self.handlers[event] = f
                return f
            return deco

        def emit(self, event, payload=None, **k):
            EMITTED.append((event, payload))

        def run(self, *a, **k):
            pass
    fsio.SocketIO = SocketIO
    fsio.emit = lambda event, payload=None, **k: EMITTED.append((event, payload))
    sys.modules["flask_socketio"] = fsio
    # ---- PyQt5
    pyqt = types.ModuleType("PyQt5")
    qtcore = types.ModuleType("PyQt5.QtCore")

    class QObject:
        def __init__(self, *a, **k):
            pass

        def deleteLater(self):
            pass

    class _Signal:
        def __init__(self):
            self.slots = []

        def connect(self, f):
            self.slots.append(f)

        def emit(self, *a):
            for f in self.slots:
                f(*a)

    class pyqtSignal:                      # class attribute -> per-instance signal
        def __init__(self, *types_):
            self.name = None

        def __set_name__(self, owner, name):
            self.name = "_sig_" + name

        def __get__(self, obj, owner):
            if obj is None:
                return self
            if self.name not in obj.__dict__:
                obj.__dict__[self.name] = _Signal()
            return obj.__dict__[self.name]

    def pyqtSlot(*a, **k):
        return lambda f: f

    class QTimer:
        def __init__(self, parent=None):
            self.timeout = _Signal()
            self.running = False
            TIMERS.append(self)

        def start(self, ms=0):
            self.running = True

        def stop(self):
            self.running = False

        def fire(self):
            if self.running:
                self.timeout.emit()

    class QMetaObject:
        @staticmethod
        def invokeMethod(obj, name, conn=None, *args):
            return getattr(obj, name)(*args)

    class Qt:
        QueuedConnection = 2

    def Q_ARG(type_, value):
        return value

    class QTime:
        @staticmethod
        def currentTime():
            return QTime()

        def msecsSinceStartOfDay(self):
            t = time.localtime()
            return int(((t.tm_hour * 60 + t.tm_min) * 60 + t.tm_sec) * 1000 + (time.time() % 1) * 1000)

    for k, v in dict(QObject=QObject, pyqtSignal=pyqtSignal, pyqtSlot=pyqtSlot, QTimer=QTimer, QMetaObject=QMetaObject,
                     Qt=Qt, Q_ARG=Q_ARG, QTime=QTime).items():
        setattr(qtcore, k, v)
    qtnet = types.ModuleType("PyQt5.QtNetwork")

    class QUdpSocket:
        def __init__(self, parent=None):
            self.readyRead = _Signal()

        def bind(self, addr, port):
            return True

        def hasPendingDatagrams(self):
            return False

        def close(self):
            pass

    class QHostAddress:
        def __init__(self, ip):
            self.ip = ip

        def toString(self):
            return self.ip
    qtnet.QUdpSocket, qtnet.QHostAddress = QUdpSocket, QHostAddress
    qtw = types.ModuleType("PyQt5.QtWidgets")
    qtw.QApplication = type("QApplication", (), {"__init__": lambda self, *a: None, "exec_": lambda self: 0, "quit": lambda self: None})
    pyqt.QtCore, pyqt.QtNetwork, pyqt.QtWidgets = qtcore, qtnet, qtw
    sys.modules.update({"PyQt5": pyqt, "PyQt5.QtCore": qtcore, "PyQt5.QtNetwork": qtnet, "PyQt5.QtWidgets": qtw})
    # ---- matplotlib (imported for the filter-preview plot only)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.use = lambda *a, **k: None
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    # ---- serial: opening a port fails, as on a machine without the board
    serial = types.ModuleType("serial")

    class SerialException(Exception):
        pass

    def Serial(*a, **k):
        raise SerialException("no serial port in the test environment")
    serial.Serial, serial.SerialException = Serial, SerialException
    sys.modules["serial"] = serial


def load_gui(source: str, name="fft_analyzer_gui_under_test"):
    """exec the (patched) GUI source as a module; its __main__ block does not run."""
    install()
    mod = types.ModuleType(name)
    mod.__file__ = name + ".py"
    sys.modules[name] = mod
    exec(compile(source, mod.__file__, "exec"), mod.__dict__)
    return mod
