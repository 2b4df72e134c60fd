"""Derive the small model-input fixtures from the reference's data/ directory.

Run in the build container (needs /root/reference); the output, reina_b200/data/inputs.json, is
committed so that tests, smoke() and bench.py never read /root/reference at run time.

Restates (no code copied):
  * calc/datasets.py:19-61  population by 1-year age for a hospital district (sum over the
    district's municipalities, both sexes).  The municipality lists come from
    data/shp_jasenkunnat_2020.xls (BIFF8; xlrd is not installed here) and are the ones recorded in
    SURVEY.md Appendix A; 'Koski tl' does not match the population file's spelling and is silently
    dropped by `isin`, exactly as in the reference.
  * calc/datasets.py:64-79  FI contact matrix: strip 'cnt_', 'otherplace'->'other', '70+'->'70-100'.
"""
import csv
import json
import os
import sys

REF = os.environ.get('REF', '/root/reference')

MUNICIPALITIES = {
    'HUS': ['Askola', 'Espoo', 'Hanko', 'Helsinki', 'Hyvinkää', 'Inkoo', 'Järvenpää', 'Karkkila',
            'Kauniainen', 'Kerava', 'Kirkkonummi', 'Lapinjärvi', 'Lohja', 'Loviisa', 'Mäntsälä',
            'Nurmijärvi', 'Pornainen', 'Porvoo', 'Raasepori', 'Sipoo', 'Siuntio', 'Tuusula',
            'Vantaa', 'Vihti'],
    'Varsinais-Suomi': ['Aura', 'Kaarina', 'Kemiönsaari', 'Koski tl', 'Kustavi', 'Laitila',
                        'Lieto', 'Loimaa', 'Marttila', 'Masku', 'Mynämäki', 'Naantali',
                        'Nousiainen', 'Oripää', 'Paimio', 'Parainen', 'Punkalaidun', 'Pyhäranta',
                        'Pöytyä', 'Raisio', 'Rusko', 'Salo', 'Sauvo', 'Somero', 'Taivassalo',
                        'Turku', 'Uusikaupunki', 'Vehmaa'],
}


def population_by_age(area):
    names = set(MUNICIPALITIES[area])
    counts = [0] * 101
    with open(os.path.join(REF, 'data/005_11re_2019.csv'), encoding='iso8859-1') as f:
        f.readline()
        f.readline()
        rd = csv.reader(f, delimiter=';', quotechar='"')
        header = next(rd)
        i_male = header.index('Miehet 2019 Väestö 31.12.')
        i_female = header.index('Naiset 2019 Väestö 31.12.')
        for row in rd:
            if not row or row[0] not in names or row[1] == 'Yhteensä':
                continue
            age = 100 if row[1].startswith('100') else int(row[1])
            counts[age] += int(row[i_male]) + int(row[i_female])
    return counts


def contact_matrix(country='FI', max_age=100):
    out = []
    with open(os.path.join(REF, 'data/contact_matrix.csv')) as f:
        rd = csv.reader(f)
        header = next(rd)
        bands = [h.replace('+', '-%d' % max_age) for h in header[3:]]
        for row in rd:
            if row[0] != country:
                continue
            place = row[1].replace('cnt_', '').replace('otherplace', 'other')
            part = row[2].replace('+', '-%d' % max_age)
            out.append(dict(place_type=place, participant_age=part,
                            contacts=[float(x) for x in row[3:]]))
    return dict(contact_bands=bands, rows=out)


def case_files():
    """data/hosp_cases_<area>.csv (calc/datasets.py:80-84): date -> [dead, in_icu, in_ward, confirmed], the inputs of
    get_initial_population_condition (:143-173)."""
    out = {}
    for area, name in (('HUS', 'hosp_cases_hus.csv'), ('Varsinais-Suomi', 'hosp_cases_varsinais-suomi.csv')):
        rows = {}
        with open(os.path.join(REF, 'data', name)) as f:
            for r in csv.DictReader(f):
                rows.setdefault(r['date'], [int(float(r[k] or 0)) for k in ('dead', 'in_icu', 'in_ward', 'confirmed')])
        out[area] = rows
    return out


def main():
    data = dict(
        areas={a: population_by_age(a) for a in MUNICIPALITIES},
        contacts=contact_matrix(),
        cases=case_files(),
    )
    for a, c in data['areas'].items():
        print(a, sum(c), file=sys.stderr)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       'reina_b200', 'data', 'inputs.json')
    with open(dst, 'w') as f:
        json.dump(data, f)
    print('wrote', dst, os.path.getsize(dst), 'bytes', file=sys.stderr)


if __name__ == '__main__':
    main()
