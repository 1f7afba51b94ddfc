"""Host-side optimizer loops over the engine's cost path: the thin layer that turns the batched
cost+gradient entry points into ``Start_Decomposition``.

Not a port of the reference's optimizers (optimization_engines/*.cpp are out of the hot-path scope and stay usable unchanged
through the drop-in of integration/): a limited-memory BFGS whose line search evaluates ALL trial step lengths of an iteration
as one batch on the device (sqgpu_line_search_batched; the reference's BFGS_Powell evaluates them one by one,
common/BFGS_Powell.cpp:70-200), and a driver for the device-resident ADAM trajectories (sqgpu_adam_steps).

Everything numerical happens in the two callables handed in -- in the product they are Engine.cost_grad_batched and
Engine.line_search_batched -- so the host logic is testable on CPU against the oracle.
"""
import numpy as np


def lbfgs(cost_grad, line_search, x0, max_iter=500, tol=1e-10, gtol=1e-9, history=12, alphas=None, callback=None):
    """Minimise f from x0. ``cost_grad(x) -> (f, g)``; ``line_search(x, d, alphas) -> (costs, dphis)`` evaluates
    f(x + a d) and d/da f(x + a d) for every a in ``alphas`` at once (one device batch).

    Per iteration: two-loop recursion for the direction, ONE batched line-search call over a geometric ladder of step lengths,
    the best point that satisfies the Armijo condition (preferring one that also satisfies the curvature condition), then one
    cost+gradient evaluation at the accepted point. ``history = 0``: steepest descent with the same line search (the reference's
    Grad_Descend, common/grad_descend.cpp:459-480). Returns (x, f, n_iterations, n_evaluations)."""
    x = np.array(x0, dtype=np.float64).reshape(-1)
    if alphas is None:
        alphas = np.array([2.0 ** k for k in range(3, -14, -1)])  # 8, 4, 2, 1, 1/2, ... 2^-13
    f, g = cost_grad(x)
    n_eval = 1
    S, Y = [], []
    it = 0
    for it in range(1, max_iter + 1):
        if f < tol or np.abs(g).max() < gtol:
            break
        # two-loop recursion
        q = g.copy()
        al = []
        for s, y in zip(reversed(S), reversed(Y)):
            a = (s @ q) / (y @ s)
            al.append(a)
            q -= a * y
        if S:
            q *= (S[-1] @ Y[-1]) / (Y[-1] @ Y[-1])
        for (s, y), a in zip(zip(S, Y), reversed(al)):
            b = (y @ q) / (y @ s)
            q += (a - b) * s
        d = -q
        dphi0 = g @ d
        if not (dphi0 < 0):  # not a descent direction: restart from steepest descent
            S, Y = [], []
            d = -g
            dphi0 = g @ d
        scale = 1.0 if S else min(1.0, 1.0 / max(np.abs(g).max(), 1e-300))
        trial = alphas * scale
        costs, dphis = line_search(x, d, trial)
        n_eval += len(trial)
        armijo = costs <= f + 1e-4 * trial * dphi0
        wolfe = armijo & (np.abs(dphis) <= 0.9 * abs(dphi0))
        pick = None
        if wolfe.any():
            idx = np.flatnonzero(wolfe)
            pick = idx[np.argmin(costs[idx])]
        elif armijo.any():
            idx = np.flatnonzero(armijo)
            pick = idx[np.argmin(costs[idx])]
        if pick is None:
            if S:  # the quasi-Newton model is bad here: drop it and retry along the gradient
                S, Y = [], []
                continue
            break  # no decrease even along the gradient at step 2^-13 / |g|: converged to rounding
        x_new = x + trial[pick] * d
        f_new, g_new = cost_grad(x_new)
        n_eval += 1
        s, y = x_new - x, g_new - g
        if s @ y > 1e-14 * np.sqrt((s @ s) * (y @ y)):
            S.append(s)
            Y.append(y)
            if len(S) > history:
                S.pop(0)
                Y.pop(0)
        x, f, g = x_new, f_new, g_new
        if callback is not None:
            callback(it, x, f)
    return x, float(f), it, n_eval


def multistart_lbfgs(cost_batched, cost_grad, line_search, n_params, rng, starts=64, keep=4, **kw):
    """``starts`` random initial points are scored with ONE batched cost call; L-BFGS runs from the ``keep`` best, best first,
    and stops at the first run that reaches kw['tol'] (the reference restarts BFGS from perturbed points one run after the
    other, optimization_engines/BFGS.cpp:124-150)."""
    X0 = rng.random((starts, n_params)) * 2 * np.pi
    X0[0] = 0.0  # the reference's other initial guess (guess_type ZEROS / CLOSE_TO_ZERO)
    scores = np.asarray(cost_batched(X0))
    best = (None, np.inf, 0, 0)
    total_eval = starts
    for i in np.argsort(scores)[:keep]:
        x, f, it, ne = lbfgs(cost_grad, line_search, X0[i], **kw)
        total_eval += ne
        if f < best[1]:
            best = (x, f, it, total_eval)
        if f < kw.get("tol", 1e-10):
            break
    return best[0], best[1], best[2], total_eval


def cosine_updates(f0, f_half, f_full, double_period=False):
    """Per-parameter step to the minimum of the sinusoid through three points (COSINE.cpp:255-291, double period :293-330):
    with f(t) = offset + A cos(w t + phi0), f0 = f(0), f_half = f(s), f_full = f(2 s) where s = pi/2 (w = 1) or pi/4 (w = 2).
    Returns the shift of t that lands on the minimum, in (-pi/w, pi/w]."""
    f_half, f_full = np.asarray(f_half, dtype=np.float64), np.asarray(f_full, dtype=np.float64)
    a_cos = (f0 - f_full) / 2
    offset = (f0 + f_full) / 2
    a_sin = offset - f_half
    phi0 = np.arctan2(a_sin, a_cos)
    if double_period:
        return np.where(phi0 > 0, np.pi / 2 - phi0 / 2, -phi0 / 2 - np.pi / 2)
    return np.where(phi0 > 0, np.pi - phi0, -phi0 - np.pi)


def cosine(cost_batched, x0, rng, batch_size=None, max_iter=2000, tol=1e-10, double_period=False, line_points=16,
           check_for_convergence=True, callback=None, cost_shifted=None):
    """The reference's COSINE engine (optimization_engines/COSINE.cpp:60-657) over a batched cost: per iteration, ``batch_size``
    distinct random parameters are each moved to the minimum of the cost as a function of that parameter alone (a sinusoid,
    fixed by the current value and two shifted evaluations), and the joint move is scaled by a line search on [0, 1].

    Same update rule and stopping rules; the evaluations are arranged for the device: the 2 x batch_size shifted parameter
    sets of an iteration are ONE cost_batched call (the reference issues two batched calls), and the line search is ONE
    batched call over ``line_points`` step fractions instead of ~12 dependent single evaluations of a golden-section search
    (COSINE.cpp:417-523; the reference keeps the batched grid variant commented out at :531-568) -- two device round trips per
    iteration instead of fourteen. ``double_period``: the VQE flavour (shifts pi/4, pi/2).

    ``cost_shifted(X, shifts) -> (cost[b], shifted[s, b, p])`` (Engine.cost_shifted_batched; decomposition costs only): the
    shifted costs of ALL parameters for both shifts come from one adjoint sweep (about three forward passes) -- it replaces the
    2 x batch_size forward passes of the shift batch, whatever the batch size. Returns (x, f, iterations, evals)."""
    x = np.array(x0, dtype=np.float64).reshape(-1)
    P = x.size
    bs = min(64, P) if batch_size is None else int(batch_size)
    if bs > P:
        raise Exception("cosine: batch size should be lower or equal to the number of free parameters")
    f = float(np.asarray(cost_batched(x.reshape(1, -1)))[0])
    n_eval = 1
    if P == 0:
        return x, f, 0, n_eval
    shift = np.pi / 4 if double_period else np.pi / 2
    fractions = np.arange(1, line_points + 1, dtype=np.float64) / line_points
    hist = np.zeros(100)
    hist_mean, hist_idx = 0.0, 0
    it = 0
    for it in range(1, max_iter + 1):
        idx = rng.choice(P, size=bs, replace=False)
        if cost_shifted is not None:
            both = np.asarray(cost_shifted(x.reshape(1, -1), (shift, 2 * shift))[1])  # [2][1][P]: one sweep serves both shifts
            f_half, f_full = both[0, 0, idx], both[1, 0, idx]
        else:
            X = np.repeat(x.reshape(1, -1), 2 * bs, axis=0)
            X[np.arange(bs), idx] += shift
            X[bs + np.arange(bs), idx] += 2 * shift
            vals = np.asarray(cost_batched(X))
            f_half, f_full = vals[:bs], vals[bs:]
        upd = cosine_updates(f, f_half, f_full, double_period)
        L = np.repeat(x.reshape(1, -1), line_points, axis=0)
        L[:, idx] += fractions[:, None] * upd[None, :]
        lv = np.asarray(cost_batched(L))
        n_eval += (1 if cost_shifted is not None else 2 * bs) + line_points
        k = int(np.argmin(lv))
        if lv[k] < f:
            x, f = L[k].copy(), float(lv[k])
        if callback is not None:
            callback(it, x, f)
        if f < tol:
            break
        hist_mean += (f - hist[hist_idx]) / hist.size
        hist[hist_idx] = f
        hist_idx = (hist_idx + 1) % hist.size
        var = np.sqrt(((hist - hist_mean) ** 2).sum()) / hist.size
        if check_for_convergence and hist_mean != 0 and abs((hist_mean - f) / hist_mean) < 1e-7 and abs(var / hist_mean) < 1e-7:
            break
    return x, f, it, n_eval


def grad_descend_shift_rule(cost_batched, x0, rng, batch_size=None, max_iter=2000, tol=1e-10, eta=1e-3, use_line_search=True,
                            line_points=128, callback=None, cost_shifted=None):
    """The reference's GRAD_DESCEND_PARAMETER_SHIFT_RULE engine (optimization_engines/GRAD_DESCEND_PARAMETER_SHIFT_RULE.cpp:
    60-400): per iteration ``batch_size`` distinct random parameters get the gradient component f(+pi/4) - f(-pi/4), scaled by
    ``eta``; the step along it is either taken as it is or (default) scaled by the best of ``line_points`` fractions k /
    line_points, evaluated as one batch (:296-325, already batched in the reference). Same update and stopping rules; the two
    shifted batches are one cost_batched call -- or, with ``cost_shifted`` (Engine.cost_shifted_batched), one adjoint sweep that
    returns both shifted costs of every parameter. Returns (x, f, iterations, evaluations)."""
    x = np.array(x0, dtype=np.float64).reshape(-1)
    P = x.size
    bs = min(64, P) if batch_size is None else int(batch_size)
    if bs > P:
        raise Exception("grad_descend_shift_rule: batch size should be lower or equal to the number of free parameters")
    f = float(np.asarray(cost_batched(x.reshape(1, -1)))[0])
    n_eval = 1
    if P == 0:
        return x, f, 0, n_eval
    fractions = np.arange(line_points, dtype=np.float64) / line_points  # includes 0: the current point (:300-306)
    hist = np.zeros(100)
    hist_mean, hist_idx = 0.0, 0
    it = 0
    for it in range(1, max_iter + 1):
        idx = rng.choice(P, size=bs, replace=False)
        if cost_shifted is not None:
            both = np.asarray(cost_shifted(x.reshape(1, -1), (np.pi / 4, -np.pi / 4))[1])
            f_plus, f_minus = both[0, 0, idx], both[1, 0, idx]
            n_eval += 1
        else:
            X = np.repeat(x.reshape(1, -1), 2 * bs, axis=0)
            X[np.arange(bs), idx] += np.pi / 4
            X[bs + np.arange(bs), idx] -= np.pi / 4
            vals = np.asarray(cost_batched(X))
            f_plus, f_minus = vals[:bs], vals[bs:]
            n_eval += 2 * bs
        upd = (f_plus - f_minus) * eta
        if use_line_search:
            L = np.repeat(x.reshape(1, -1), line_points, axis=0)
            L[:, idx] -= fractions[:, None] * upd[None, :]
            lv = np.asarray(cost_batched(L))
            n_eval += line_points
            k = int(np.argmin(lv))
            x, f = L[k].copy(), float(lv[k])
        else:
            x[idx] -= upd
            f = float(np.asarray(cost_batched(x.reshape(1, -1)))[0])
            n_eval += 1
        if callback is not None:
            callback(it, x, f)
        if f < tol:
            break
        hist_mean += (f - hist[hist_idx]) / hist.size
        hist[hist_idx] = f
        hist_idx = (hist_idx + 1) % hist.size
        var = np.sqrt(((hist - hist_mean) ** 2).sum()) / hist.size
        if hist_mean != 0 and abs(hist_mean - f) < 1e-7 and var / hist_mean < 1e-7:
            break
    return x, f, it, n_eval


def five_point_updates(theta, f0, f_pi4, f_pi2, f_pi, f_3pi2):
    """The five-point rule of AGENTS for costs quadratic in the trace (Hilbert-Schmidt test; AGENTS.cpp:583-660): along one
    parameter f(p) = kappa sin(2 p + xi) + gamma sin(p + varphi) + offset, fixed by the value at p = theta and at the shifts
    pi/4, pi/2, pi, 3 pi/2. Returns (shift to the minimum of that curve, its value there). The reference minimises the curve
    with ten BFGS_Powell iterations from an analytic guess; here: a 1440-point scan of the period, then Newton steps."""
    theta, f0, f_pi4, f_pi2, f_pi, f_3pi2 = (np.asarray(a, dtype=np.float64) for a in (theta, f0, f_pi4, f_pi2, f_pi, f_3pi2))
    f1 = f0 - f_pi
    f2 = f_pi2 - f_3pi2
    gamma = 0.5 * np.sqrt(f1 * f1 + f2 * f2)
    varphi = np.arctan2(f1, f2) - theta
    offset = 0.25 * (f0 + f_pi + f_pi2 + f_3pi2)
    f3 = 0.5 * (f0 + f_pi - 2 * offset)
    f4 = f_pi4 - offset - gamma * np.sin(theta + np.pi / 4 + varphi)
    kappa = np.sqrt(f3 * f3 + f4 * f4)
    xi = np.arctan2(f3, f4) - 2 * theta
    curve = lambda p: kappa * np.sin(2 * p + xi) + gamma * np.sin(p + varphi) + offset
    grid = np.linspace(0.0, 2 * np.pi, 1440, endpoint=False)
    vals = curve(grid[:, None]) if theta.ndim else curve(grid)
    p = grid[np.argmin(vals, axis=0)]
    for _ in range(4):
        d1 = 2 * kappa * np.cos(2 * p + xi) + gamma * np.cos(p + varphi)
        d2 = -4 * kappa * np.sin(2 * p + xi) - gamma * np.sin(p + varphi)
        step = np.where(d2 > 1e-300, d1 / np.where(d2 > 1e-300, d2, 1.0), 0.0)
        p = p - np.clip(step, -0.01, 0.01)
    shift = np.mod(p - theta + np.pi, 2 * np.pi) - np.pi
    return shift, curve(p)


def agents(cost_batched, x0, rng, agent_num=64, max_iter=10000, tol=1e-10, double_period=False, agent_lifetime=1000,
           exploration_rate=0.2, agent_randomization_rate=0.2, randomization_rate=0.3, radius=1.0, convergence_length=20,
           scale_by_cost=True, callback=None, five_point=False):
    """The reference's AGENTS engine (optimization_engines/AGENTS.cpp:60-938), three-point rule (Frobenius cost; doubled period for
    the VQE energy, :333-335): ``agent_num`` parameter vectors walk independently -- per iteration every agent draws ONE of its
    parameters and jumps to the minimum of the sinusoid the cost is along it (exact, no line search: the new cost is offset -
    amplitude, :395-415) -- and every ``agent_lifetime`` iterations the costs are re-evaluated, the best agent is recorded,
    and each worse agent adopts the best one's state with probability ``exploration_rate``, then perturbs it with probability
    ``agent_randomization_rate`` (randomize_parameters, Optimization_Interface.cpp:576-603: each parameter with probability
    ``randomization_rate`` moves by U(-2 pi, 2 pi) * radius [* sqrt(f) for the decomposition costs]).

    Arranged for the device: the two shifted parameter sets of all agents are ONE cost_batched call of 2 x agent_num rows per
    iteration (the reference issues two), and the perturbed agents of a census are re-scored by one batched call instead of one
    single evaluation each. Stops when an agent is below ``tol`` or the recorded minimum stalls over ``convergence_length``
    censuses (:849-866). ``five_point``: the rule for costs quadratic in the trace (Hilbert-Schmidt test, :335, :536-660): four
    shifted sets per agent instead of two, still one batch per iteration. Returns (x, f, iterations, evaluations)."""
    x0 = np.array(x0, dtype=np.float64).reshape(-1)
    P = x0.size
    if P == 0:
        return x0, float(np.asarray(cost_batched(x0.reshape(1, -1)))[0]), 0, 1
    A = int(agent_num)
    shift = np.pi / 4 if double_period else np.pi / 2

    def randomized(src, f_ref):
        out = src.copy()
        mask = rng.random(P) <= randomization_rate
        step = rng.uniform(-2 * np.pi, 2 * np.pi, P) * radius * (np.sqrt(max(f_ref, 0.0)) if scale_by_cost else 1.0)
        out[mask] += step[mask]
        return out

    f_best = float(np.asarray(cost_batched(x0.reshape(1, -1)))[0])
    x_best = x0.copy()
    X = np.stack([x0] + [randomized(x0, f_best) for _ in range(A - 1)])
    F = np.asarray(cost_batched(X), dtype=np.float64).copy()
    n_eval = 1 + A
    hist = np.zeros(convergence_length)
    hist_mean, hist_idx = 0.0, 0
    rows = np.arange(A)
    it = 0
    for it in range(max_iter):
        idx = rng.integers(0, P, A)
        if five_point:
            S = np.concatenate([X, X, X, X])
            for k, sh in enumerate((np.pi / 4, np.pi / 2, np.pi, 3 * np.pi / 2)):
                S[k * A + rows, idx] += sh
            vals = np.asarray(cost_batched(S))
            n_eval += 4 * A
            upd, F = five_point_updates(X[rows, idx], F, vals[:A], vals[A:2 * A], vals[2 * A:3 * A], vals[3 * A:])
            X[rows, idx] += upd
        else:
            S = np.concatenate([X, X])
            S[rows, idx] += shift
            S[A + rows, idx] += 2 * shift
            vals = np.asarray(cost_batched(S))
            n_eval += 2 * A
            f_half, f_full = vals[:A], vals[A:]
            a_cos, offset = (F - f_full) / 2, (F + f_full) / 2
            a_sin = offset - f_half
            X[rows, idx] += cosine_updates(F, f_half, f_full, double_period)
            F = offset - np.sqrt(a_sin * a_sin + a_cos * a_cos)
        census = it % agent_lifetime == 0
        if census:
            F = np.asarray(cost_batched(X), dtype=np.float64).copy()  # the predicted costs drift by rounding: re-evaluate (:700-705)
            n_eval += A
            b = int(np.argmin(F))
            if F[b] <= f_best:
                f_best, x_best = float(F[b]), X[b].copy()
            moved = []
            for a in range(A):
                if a != b and F[a] > f_best and rng.random() < exploration_rate:
                    X[a], F[a] = X[b], F[b]
                    if rng.random() < agent_randomization_rate:
                        X[a] = randomized(x_best, f_best)
                        moved.append(a)
            if moved:
                F[moved] = np.asarray(cost_batched(X[moved]))
                n_eval += len(moved)
            hist_mean += (f_best - hist[hist_idx]) / hist.size
            hist[hist_idx] = f_best
            hist_idx = (hist_idx + 1) % hist.size
            var = np.sqrt(((hist - hist_mean) ** 2).sum()) / hist.size
            if callback is not None:
                callback(it, x_best, f_best)
            if abs(hist_mean - f_best) < 1e-7 and var < 1e-7:
                break
        if F.min() < tol:
            b = int(np.argmin(F))
            f_chk = float(np.asarray(cost_batched(X[b:b + 1]))[0])  # the stop is decided on an evaluated cost, not a predicted one
            n_eval += 1
            F[b] = f_chk
            if f_chk < f_best:
                f_best, x_best = f_chk, X[b].copy()
            if f_chk < tol:
                break
    b = int(np.argmin(F))
    if F[b] < f_best:
        f_chk = float(np.asarray(cost_batched(X[b:b + 1]))[0])
        n_eval += 1
        if f_chk < f_best:
            f_best, x_best = f_chk, X[b].copy()
    return x_best, f_best, it + 1, n_eval
