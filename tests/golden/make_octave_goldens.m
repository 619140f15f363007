function make_octave_goldens(refroot, indir, outdir)
% make_octave_goldens(refroot, indir, outdir)
%
% Pins the hot-path oracle to the REFERENCE ITSELF.  Runs the UNMODIFIED functions of ndefilippis/SWRaytracing (checked
% out at `refroot`, default /root/reference) under MATLAB or GNU Octave on the seeded inputs of
% tests/golden/hotpath_nx32.npz (exported to `indir` = tests/golden/octave_in by export_hotpath_inputs.py) and writes every
% result to `outdir` (default tests/golden/octave_out) with the reference's own write_field.m (raw native-endian real*8,
% column-major).  tests/test_octave_goldens.py picks the files up: when they exist, the CPU oracle AND the GPU path are
% held to them (fields / RHS 1e-12, trajectories 1e-9); until someone runs this script the hot-path parity stays "unpinned".
%
% Usage (from the repo root, on any machine that has the reference checkout and MATLAB >= R2019b or Octave >= 7):
%     octave --eval "addpath('tests/golden'); make_octave_goldens('/root/reference')"
%     matlab -batch "addpath('tests/golden'); make_octave_goldens('/root/reference')"
% then commit tests/golden/octave_out/*.bin.
%
% (The committed octave_out/ was produced by tests/golden/run_reference_recipe.py: this very file and the reference's .m files
% executed by oracle/minimat, the MATLAB-subset interpreter of this repository, because the build image has no MATLAB / Octave.)
%
% The reference holds TWO copies of interpolate.m that differ in one line: ray_trace_sw/interpolate.m:13 (bump 1e-13) and
% qg_flow_ray_trace/interpolate.m:13 (bump 1e-10).  Which one a caller binds to is a matter of the MATLAB path, and this recipe
% sets the path the way each caller meets it in the reference's own runs: interpolate_U and the ode23 right-hand side run with
% qg_flow_ray_trace in front (runqgsw_raytrace.sbatch:25-27 copies that folder's interpolate.m next to qgsw_raytrace.m), everything
% else with ray_trace_sw in front (SpectralScheme.m:8 does that itself; raytrace.m / raytrace_sw.m live in that folder).
%
% Reference functions exercised (file:line = what the oracle restates):
%   interpolate            ray_trace_sw/interpolate.m:1-50, qg_flow_ray_trace/interpolate.m:1-50
%   interpolate_U          qg_flow_ray_trace/interpolate_U.m:1-24
%   SpectralScheme         SpectralScheme.m:6-36 (ctor: g2k, ik-multiplication, k2g), :45-54 (U), :56-68 (grad_U)
%   grad_U_times_k         RaytracingScheme.m:9-16
%   ode_symplectic         ode_symplectic.m:1-37 (100 leapfrog steps)
%   cg_sw                  ray_trace_sw/cg_sw.m:15-31
%   step_packet            ray_trace_sw/step_packet.m:37-78 (3 steps per packet)
%   step_packet_xka        ray_trace_sw/step_packet_xka.m:38-91 (3 steps per packet)
if nargin < 1 || isempty(refroot), refroot = '/root/reference'; end
here = fileparts(mfilename('fullpath'));
if nargin < 2 || isempty(indir), indir = fullfile(here, 'octave_in'); end
if nargin < 3 || isempty(outdir), outdir = fullfile(here, 'octave_out'); end
if ~exist(outdir, 'dir'), mkdir(outdir); end
old = dir(fullfile(outdir, '*.bin'));
for i = 1:numel(old), delete(fullfile(outdir, old(i).name)); end      % write_field APPENDS (fopen 'a', write_field.m:31)

startdir = pwd;
cd(refroot);                                   % SpectralScheme's constructor does addpath ./rsw/ ./ray_trace_sw/
addpath(refroot);                                  % SpectralScheme, RaytracingScheme, ode_symplectic
addpath(fullfile(refroot, 'ray_trace_sw'));        % interpolate (bump 1e-13), step_packet, step_packet_xka, cg_sw
addpath(fullfile(refroot, 'qg_flow_ray_trace'));   % read_field, write_field, interpolate_U, interpolate (bump 1e-10); now in FRONT
rd = @(name, m, n) read_field(fullfile(indir, name), m, n, 1, 1, 1);
put = @(name, A) write_field(A, fullfile(outdir, name), 1);

p = rd('params', 8, 1);
nx = p(1); L = p(2); f = p(3); gH = p(4); alpha = p(5); dt = p(6); C0 = p(7); np_ = p(8);
dx = L / nx;
x = rd('x', np_, 1); y = rd('y', np_, 1); k = rd('k', np_, 1); l = rd('l', np_, 1);
x = x(:); y = y(:); k = k(:); l = l(:);
names = {'u', 'v', 'ux', 'uy', 'vx', 'vy'};
for c = 1:6
    bf1.(names{c}) = rd(['bf1_' names{c}], nx, nx);
    bf2.(names{c}) = rd(['bf2_' names{c}], nx, nx);
end
H = rd('H', nx, nx);
psi = rd('psi', nx, nx);

% ==== context of the QG drivers: qg_flow_ray_trace is the first folder of the path, its interpolate.m (bump 1e-10) is the one
%      interpolate_U and the ode23 right-hand side bind to ======================================================================
E = zeros(6, np_);
for c = 1:6, E(c, :) = interpolate(x, y, bf1.(names{c}), dx, dx); end
put('eval_lagrange_qg', E);

% ---- interpolate_U: two frames blended at alpha -------------------------------------------------------------------
[U, nab] = interpolate_U(bf1, bf2, alpha, [x y], dx);
put('interpU_lagrange', [U(:, 1)'; U(:, 2)'; nab.u_x(:)'; nab.u_y(:)'; nab.v_x(:)'; nab.v_y(:)']);

% ---- the ode23 right-hand side of qgsw_raytrace.m:259-265 (Cg = 1): dx/dt = U + Cg k/omega, dk/dt = -(grad U)^T k --------
Cg = 1;
om = sqrt(f^2 + Cg^2 * (k.^2 + l.^2));
dxdt = U + Cg * [k l] ./ om;
dkdt = -[nab.u_x .* k + nab.v_x .* l, nab.u_y .* k + nab.v_y .* l];
put('rhs_lagrange', [dxdt(:, 1)'; dxdt(:, 2)'; dkdt(:, 1)'; dkdt(:, 2)']);

% ==== context of SpectralScheme / step_packet* / raytrace*: ray_trace_sw in front (addpath moves it there; SpectralScheme.m:8
%      does the same), interpolate.m with bump 1e-13 =============================================================================
addpath(fullfile(refroot, 'ray_trace_sw'));
E = zeros(6, np_);
for c = 1:6, E(c, :) = interpolate(x, y, bf1.(names{c}), dx, dx); end
put('eval_lagrange', E);

% ---- SpectralScheme: constructor from the streamfunction grid, U, grad_U, grad_U_times_k ----------------------------
scheme = SpectralScheme(L, nx, psi);
x3 = zeros(1, 2, np_); x3(1, 1, :) = x; x3(1, 2, :) = y;
k3 = zeros(1, 2, np_); k3(1, 1, :) = k; k3(1, 2, :) = l;
Us = scheme.U(x3, 0);
g = scheme.grad_U(x3, 0);
gk = scheme.grad_U_times_k(x3, k3, 0);
put('scheme_eval', [squeeze(Us(1, 1, :))'; squeeze(Us(1, 2, :))'; g.u_x(:)'; g.u_y(:)'; g.v_x(:)'; g.v_y(:)']);
put('scheme_gradU_times_k', [squeeze(gk(1, 1, :))'; squeeze(gk(1, 2, :))']);
put('scheme_fields', cat(3, scheme.U_field.u, scheme.U_field.v, scheme.GradU_field.u_x, scheme.GradU_field.u_y, ...
                         scheme.GradU_field.v_x, scheme.GradU_field.v_y));

% ---- ode_symplectic: 100 leapfrog steps (Nsteps = floor(T/dt) rows, Nsteps-1 steps; row 1 = initial state) -----------
nst = 100;
[xs, ks, ts] = ode_symplectic(x3, k3, dt, (nst + 1.5) * dt, f, gH, scheme);
assert(size(xs, 1) == nst + 1);
put('leapfrog100_scheme', [squeeze(xs(end, 1, :))'; squeeze(xs(end, 2, :))'; squeeze(ks(end, 1, :))'; squeeze(ks(end, 2, :))']);
put('leapfrog20_scheme', [squeeze(xs(21, 1, :))'; squeeze(xs(21, 2, :))'; squeeze(ks(21, 1, :))'; squeeze(ks(21, 2, :))']);
put('leapfrog_t', ts(:));

% ---- cg_sw on the first packet: whole-grid omega, C, div C, grad omega ----------------------------------------------
Uf.u = bf1.u; Uf.v = bf1.v;
Gf.u_x = bf1.ux; Gf.u_y = bf1.uy; Gf.v_x = bf1.vx; Gf.v_y = bf1.vy;
[C, omg, ~, divC, gom] = cg_sw(k(1), l(1), C0, f, Uf, H);
put('cg_sw_fields', cat(3, C.x, C.y, omg, divC, gom.x, gom.y));

% ---- step_packet / step_packet_xka: three RK4 steps per packet ---------------------------------------------------------
S1 = zeros(4, np_); S2 = zeros(5, np_);
for m = 1:np_
    P.x = x(m); P.y = y(m); P.k = k(m); P.l = l(m);
    Q = P; Q.a = 1;
    for s = 1:3
        P = step_packet(P, Uf, Gf, C0, f, dx, dx, dt);
        Q = step_packet_xka(Q, Uf, Gf, H, C0, f, dx, dx, dt);
    end
    S1(:, m) = [P.x; P.y; P.k; P.l];
    S2(:, m) = [Q.x; Q.y; Q.k; Q.l; Q.a];
end
put('rk4x3_packet_lagrange', S1);
put('rk4x3_xka_lagrange', S2);

cd(startdir);
fprintf('make_octave_goldens: wrote %d files to %s\n', numel(dir(fullfile(outdir, '*.bin'))), outdir);
end
