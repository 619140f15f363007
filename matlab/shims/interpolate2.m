function FI = interpolate2(x, y, F, dx, dy)
% interpolate2.m in the reference is an abandoned interp2(...,'cubic') experiment with no live
% call site; the name is kept as an alias of interpolate.
FI = interpolate(x, y, F, dx, dy);
end
