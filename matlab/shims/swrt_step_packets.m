function Pout = swrt_step_packets(P, U, GradU, H, C0, f, dx, dt, scheme)
% shared body of step_packet / step_packet_xka: flow uploaded once per (U, H) and kept on the device
persistent eng key
nx = size(U.u, 1);
k = [sum(U.u(:)), sum(U.v(:)), nx, dx, C0, f, ~isempty(H)];
if isempty(eng) || ~isequal(k, key)
    if ~isempty(eng), swrt_mex('destroy', eng); end
    eng = swrt_mex('create', nx, dx * nx, f, C0^2, 1);                  % LAGRANGE6 = reference semantics
    if isempty(H)
        swrt_mex('set_flow_grid', eng, 0, U.u, U.v, GradU.u_x, GradU.u_y, GradU.v_x, GradU.v_y);
    else
        swrt_mex('set_flow_grid', eng, 0, U.u, U.v, GradU.u_x, GradU.u_y, GradU.v_x, GradU.v_y, H);
    end
    key = k;
end
if isfield(P, 'a'), a = [P.a]; else, a = ones(1, numel(P)); end
swrt_mex('set_packets', eng, [P.x], [P.y], [P.k], [P.l], a);
swrt_mex('step', eng, scheme, dt, 1);
[x, y, kk, ll, a] = swrt_mex('get_packets', eng);
Pout = P;
for i = 1:numel(P)
    Pout(i).x = x(i); Pout(i).y = y(i); Pout(i).k = kk(i); Pout(i).l = ll(i);
    if scheme == 2, Pout(i).a = a(i); end
end
end
