"""Parameter-record bookkeeping shared by Population and NoiseModel.

A parameter record is the reference's dict {'value', 'fixed', 'bounds', 'constraint_expr'}
(definitions.py:130).  Both owners keep exactly the records that are valid for their current
configuration; the rules are the same (reference system/population.py:63-82, system/noise.py:25-48),
so they live here once."""
import copy


def reconcile(owned, valid, updates, unknown_message):
    """Bring the mapping `owned` in line with the template `valid` and apply `updates`.

    * a name in `updates` that `valid` does not know raises ValueError(unknown_message(name));
    * records `owned` already holds survive where still valid (their fields win over the defaults),
      the rest are dropped;
    * `updates` are merged field-wise on top;
    * names keep their first-seen order in `owned`, new ones are appended in template order.
    """
    for name in updates:
        if name not in valid:
            raise ValueError(unknown_message(name))
    for name in [n for n in owned if n not in valid]:
        del owned[name]
    for name, template in valid.items():
        record = template
        if name in owned:
            record.update(owned[name])
        if name in updates:
            record.update(updates[name])
        owned[name] = record
    return owned


def snapshot(records):
    """Deep copy of a record mapping as a plain dict (for to_dict())."""
    return dict((name, copy.deepcopy(rec)) for name, rec in records.items())
