"""Caller of the hot path: the adaptive time loop of Acme2CylSimulation::run
(dune/ax1/acme2_cyl/common/acme2_cyl_simulation.hh:236-932), restated for driving the device path
(or, in tests, the CPU oracle -- both expose set_time / update_state / set_uold / newton /
accept_state / membrane_potential).

  dt control   :300-343  x1.1 if the last Newton solve took < lowerLimit iterations and not more than the
                         one before, /1.2 if it took >= upperLimit; no first-step exception (diagInfo.iterations
                         starts at 0, so the first adaptive step is already 1.1 dt0)
  equilibration :241-263 dt held while doEquilibration and time < tEquilibrium; from the first time point
                         >= tEquilibrium on, dt = dtstart on every step (the reference's equilibrationOver flag
                         is only ever set, never cleared)
  clamps       :362-363  [minTimeStep, maxTimeStep] (INI values in seconds, loop works in micro seconds)
  AP limiter   :378-383  dt <= maxTimeStepAP while stimulating or while max V_m > -50 mV
  stim start   :262-266  dt = dtstart on the step that crosses t_inj_start
  restart      :593-638  Newton failure -> dt /= 2, updateState and Newton again from the iterate the failed solve
                         left behind (:563-566), up to maxNumberNewtonRestarts; from uold after a NaN defect or
                         with restart_from_uold
  protocol     :457, 589, 916-927  updateState(time,dt) -> solver.apply -> uold = unew -> acceptTimeStep
This is SURVEY 8(f) item 2 ("next"): host control flow around the device-resident Newton solve.
"""
import numpy as np


class NewtonFailure(RuntimeError):
    pass


class TimeLoop:
    def __init__(self, dev, u0, dtstart=1.0, time0=0.0, newton_opts=None, do_stimulation=False, tinj_start=20.0,
                 tinj_end=2.0e3, max_dt=50.0, min_dt=0.05, max_dt_ap=10.0, lower_limit=3, upper_limit=5,
                 max_restarts=3, adaptive=True, do_equilibration=None, t_equilibrium=None, restart_from_uold=False):
        self.dev, self.u = dev, np.array(u0, dtype=np.float64)
        prm = getattr(dev, "params", None) or getattr(dev, "cfg", None)   # device handle / oracle problem
        if do_equilibration is None:
            do_equilibration = bool(getattr(prm, "do_equilibration", 0))
        if t_equilibrium is None:
            t_equilibrium = float(getattr(prm, "t_equilibrium", 0.0))
        self.do_equil, self.t_equil, self.restart_from_uold = do_equilibration, t_equilibrium, restart_from_uold
        self.time, self.dt, self.dtstart = float(time0), float(dtstart), float(dtstart)
        self.opts = newton_opts
        self.do_stim, self.tinj_start, self.tinj_end = do_stimulation, tinj_start, tinj_end
        self.max_dt, self.min_dt, self.max_dt_ap = max_dt, min_dt, max_dt_ap
        self.lower, self.upper, self.max_restarts, self.adaptive = lower_limit, upper_limit, max_restarts, adaptive
        self.its, self.last_its = 0, 0
        self.trace = []  # (time, dt, newton iterations, restarts, min V_m, max V_m)

    def _choose_dt(self):
        accept = not self.adaptive or (self.do_equil and self.time < self.t_equil)
        if self.do_equil and (self.time > self.t_equil or abs(self.time - self.t_equil) < 1e-6):
            self.dt = self.dtstart
            accept = True
        if self.do_stim and (self.time + self.dt > self.tinj_start and self.time < self.tinj_start):
            accept = True
            self.dt = self.dtstart
        if not accept:
            if self.its < self.lower and self.its <= self.last_its:
                self.dt *= 1.1
            if self.its >= self.upper:
                self.dt /= 1.2
            self.dt = min(self.dt, self.max_dt)
            self.dt = max(self.dt, self.min_dt)
            vm = self.dev.membrane_potential(self.u)
            stim_now = self.do_stim and self.tinj_start < self.time + self.dt < self.tinj_end
            if stim_now or vm.max() > -50.0:
                self.dt = min(self.dt, self.max_dt_ap)

    def step(self):
        self._choose_dt()
        restarts = 0
        uold = self.u
        ustart = uold
        while True:
            self.dev.set_time(self.time, self.dt)
            self.dev.update_state(ustart)        # gfMembFlux.updateState(time, dt) reads the *new* solution vector
            self.dev.set_uold(uold)              # OneStepMethod: r0 = -M(uold)
            unew, res = self.dev.newton(ustart, self.opts)
            if res.status == 0:
                break
            ustart = uold if (self.restart_from_uold or res.status == 3) else unew
            restarts += 1
            if restarts > self.max_restarts:
                raise NewtonFailure("Newton failed %d times at t=%g (status %d)" % (restarts, self.time, res.status))
            self.dt *= 0.5
        self.last_its, self.its = self.its, res.iterations
        self.time += self.dt
        self.u = unew
        self.dev.accept_state()                  # gfMembFlux.acceptTimeStep()
        vm = self.dev.membrane_potential(self.u)
        self.trace.append((self.time, self.dt, res.iterations, restarts, float(vm.min()), float(vm.max())))
        return res

    def run(self, tend, max_steps=10 ** 9, vm_every=None):
        vms = []
        n = 0
        while self.time < tend - 1e-8 and n < max_steps:
            self.step()
            n += 1
            if vm_every and n % vm_every == 0:
                vms.append((self.time, self.dev.membrane_potential(self.u)))
        return vms


class DeviceTimeLoop:
    """The same loop run by the library itself (ax1gpu_time_step, csrc/timeloop.cu): iterate, previous step and
    channel state stay in HBM, only the step result (Newton statistics, min / max V_m) comes back per step."""

    def __init__(self, dev, u0, dtstart=1.0, time0=0.0, newton_opts=None, max_dt=50.0, min_dt=0.05, max_dt_ap=10.0,
                 lower_limit=3, upper_limit=5, max_restarts=3, adaptive=True, restart_from_uold=False):
        from . import api
        self.dev = dev
        self.opts = newton_opts or api.default_newton_opts()
        self.tl = api.TimeLoopOpts()
        api.lib().ax1gpu_timeloop_opts_default(self.tl)
        self.tl.dtstart, self.tl.max_dt, self.tl.min_dt, self.tl.max_dt_ap = dtstart, max_dt, min_dt, max_dt_ap
        self.tl.lower_limit, self.tl.upper_limit = lower_limit, upper_limit
        self.tl.max_restarts, self.tl.adaptive = max_restarts, int(adaptive)
        self.tl.restart_from_uold = int(restart_from_uold)
        self.state = api.TimeLoopState(time=float(time0), dt=float(dtstart), its=0, last_its=0, steps=0)
        dev.upload_u(u0)
        self.trace = []

    @property
    def time(self):
        return self.state.time

    @property
    def dt(self):
        return self.state.dt

    @property
    def u(self):
        return self.dev.download_u()

    def step(self):
        res = self.dev.time_step(self.tl, self.opts, self.state)
        if res.newton_status != 0:
            raise NewtonFailure("Newton failed %d times at t=%g (status %d)" % (res.restarts, res.time,
                                                                               res.newton_status))
        self.trace.append((res.time, res.dt, res.newton_iterations, res.restarts, res.vm_min, res.vm_max))
        return res

    def run(self, tend, max_steps=10 ** 9):
        n = 0
        while self.state.time < tend - 1e-8 and n < max_steps:
            self.step()
            n += 1
        return n
