function Pout = swrt_step_packets(P, U, GradU, H, C0, f, dx, dt, scheme)
% shared body of step_packet / step_packet_xka: the flow is uploaded once and kept on the device for as long as the
% caller passes THE SAME fields.  "The same" is decided on the full contents of every uploaded array (isequal on U, GradU
% and H against the copies kept here -- MATLAB / Octave arrays are copy-on-write, so keeping them costs nothing until the
% caller modifies them), never on a checksum: velocity fields have zero mean, and a changed GradU or H with an unchanged U
% must not reuse a stale flow.
persistent eng cU cG cH cpar
nx = size(U.u, 1);
par = [nx, dx, C0, f];
if isempty(eng) || ~isequal(par, cpar) || ~isequal(U, cU) || ~isequal(GradU, cG) || ~isequal(H, cH)
    if ~isempty(eng), swrt_mex('destroy', eng); eng = []; end
    eng = swrt_mex('create', nx, dx * nx, f, C0^2, 1);                  % LAGRANGE6 = reference semantics
    if isempty(H)
        swrt_mex('set_flow_grid', eng, 0, U.u, U.v, GradU.u_x, GradU.u_y, GradU.v_x, GradU.v_y);
    else
        swrt_mex('set_flow_grid', eng, 0, U.u, U.v, GradU.u_x, GradU.u_y, GradU.v_x, GradU.v_y, H);
    end
    cU = U; cG = GradU; cH = H; cpar = par;
end
if isfield(P, 'a'), a = [P.a]; else, a = ones(1, numel(P)); end
% host arrays in, one step, host arrays out: ONE gateway call (swrt_step_host)
[x, y, kk, ll, a] = swrt_mex('step_host', eng, scheme, dt, 1, [P.x], [P.y], [P.k], [P.l], a);
Pout = P;
for i = 1:numel(P)
    Pout(i).x = x(i); Pout(i).y = y(i); Pout(i).k = kk(i); Pout(i).l = ll(i);
    if scheme == 2, Pout(i).a = a(i); end
end
end
